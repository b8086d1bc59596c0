"""B200-native joint CTC/attention(+RNNLM) beam-search decode path."""
