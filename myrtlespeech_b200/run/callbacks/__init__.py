from .mixed_precision import BF16MixedPrecision, ClipGradNorm  # noqa: F401
from .rnn_t_training import RNNTTraining, ReportRNNTDecoder  # noqa: F401
