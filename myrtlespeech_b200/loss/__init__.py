from .rnn_t_loss import RNNTLoss  # noqa: F401
