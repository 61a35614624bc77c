"""Drop-in for nets/teacher_training/e2e_tts_tacotron2_sa.py (reference) -- FCL-taco2-T."""
from fcl_taco2_b200.model import Tacotron2_sa  # noqa: F401
