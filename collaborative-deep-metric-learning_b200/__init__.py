"""cdml_b200 -- B200-native hot path of Collaborative Deep Metric Learning.

Host side (Python, mirroring the reference's module names) over libcdml.so (hand-written sm_100a CUDA behind a C ABI).
Import as ``cdml_b200`` (see /cdml_b200.py) -- the directory name is not a valid Python identifier."""
__version__ = "0.1.0"
