"""Drop-in for the reference's ``src/methods/zero_shot/hard_em_dirichlet.py`` (``src/eval_zero_shot.py:13``)."""
from tclip_b200.methods.dirichlet import BASE, HARD_EM_DIRICHLET  # noqa: F401
