from .rnn_t import RNNT, JointHandle, RNNTJoint, RNNTPredictionNet  # noqa: F401
