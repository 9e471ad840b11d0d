"""sassy_b200 -- B200-native approximate string matching (the Searcher::search path of Sassy).

The compute path is libsassy_b200.so (CUDA, sm_100a); importing this package never
falls back to a CPU implementation."""
from .searcher import (DeviceText, EncodedPatterns, Match, MatchList, Searcher, device_count, host_alloc,
                       host_free)

__all__ = ["Searcher", "Match", "MatchList", "DeviceText", "EncodedPatterns", "device_count", "host_alloc", "host_free"]
