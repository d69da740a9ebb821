"""See ../update/ggnn_update.py."""
