"""Drop-in for the reference's ``src/methods/few_shot/hard_em_dirichlet.py`` (``src/eval_few_shot.py:13``)."""
from tclip_b200.methods.dirichlet import FEW_SHOT_BASE as BASE  # noqa: F401
from tclip_b200.methods.dirichlet import FEW_SHOT_HARD_EM_DIRICHLET as HARD_EM_DIRICHLET  # noqa: F401
