"""Train-loop glue for the RNN-T path; the loop itself stays the reference's ``run/train.py``."""
