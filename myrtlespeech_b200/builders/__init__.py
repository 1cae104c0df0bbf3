from . import rnn_t, rnn_t_greedy_decoder, rnn_t_loss, speech_to_text  # noqa: F401
