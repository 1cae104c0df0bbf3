from .rnn_t_greedy_decoder import RNNTGreedyDecoder  # noqa: F401
