"""Drop-in for nets/knowledge_distillation/e2e_tts_tacotron2_sa_kd_student.py (reference) -- FCL-taco2-S."""
from fcl_taco2_b200.model import Tacotron2_sa_student as Tacotron2_sa  # noqa: F401
