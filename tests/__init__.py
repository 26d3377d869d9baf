"""Test suite of cap2det_b200: oracle pins (CPU), host logic (CPU), CUDA parity (gpu marker)."""
