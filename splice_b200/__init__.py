"""splice_b200 — B200-native (sm_100a) hot path of omerbt/Splice behind the reference's Python entry points."""
__version__ = "0.1.0"
