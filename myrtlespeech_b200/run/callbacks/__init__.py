from .rnn_t_training import RNNTTraining, ReportRNNTDecoder  # noqa: F401
