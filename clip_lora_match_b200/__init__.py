"""B200-native retrieval hot path of youngalip/clip-lora-match.

Host code mirrors the reference's Python surface (models/clip_model.py, models/lora_adapter.py,
src/embedding/*.py, scripts/build_*_index.py, scripts/demo_search_*); the arithmetic runs in
hand-written sm_100a kernels behind the C-ABI of include/clm_b200.h (csrc/).
"""
__version__ = "0.1.0"
