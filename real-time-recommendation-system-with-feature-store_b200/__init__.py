"""b200rec — B200-native two-tower hot path (towers, contrastive losses, exact inner-product top-K).

Python host code mirrors the reference's interfaces (src/models/two_tower.py, src/serving/retrieval.py); all
arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI declared in include/b200rec.h.
There is no CPU fallback: importing the compute modules without libb200rec.so raises.
"""
__version__ = "0.1.0"
