"""piccolo_b200 — B200-native sampling-loss pose search (drop-in for the hot path of 82magnolia/piccolo).

Host side is Python/PyTorch (plumbing); the arithmetic runs in libpiccolo_b200.so (hand-written
sm_100a CUDA behind the C ABI of include/piccolo_b200.h).  Importing the package does not need a GPU;
every compute call does, and fails loudly without one.
"""
__version__ = "0.1.0"
