"""Drop-in for the reference's ``src/methods/few_shot/em_dirichlet.py`` (imported by ``src/eval_few_shot.py:12``):
``run_task(task_dic, shot)`` with support + query; the arithmetic runs in libtclip_b200 (sm_100a CUDA)."""
from tclip_b200.methods.dirichlet import FEW_SHOT_BASE as BASE  # noqa: F401
from tclip_b200.methods.dirichlet import FEW_SHOT_EM_DIRICHLET as EM_DIRICHLET  # noqa: F401
