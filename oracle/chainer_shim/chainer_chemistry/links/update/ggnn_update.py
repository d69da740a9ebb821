"""Only imported by name in models/ggnn_gwm.py-style files; the hot path uses the reference's own models/update/ggnn_update.py."""
