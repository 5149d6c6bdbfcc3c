"""Drop-in for the reference's ``src/methods/zero_shot/em_dirichlet.py`` (imported by ``src/eval_zero_shot.py:12``):
same class names, constructor ``(model, device, log_file, args)`` and ``run_task(task_dic)``; the arithmetic runs in
libtclip_b200 (sm_100a CUDA).  See INTEGRATION.md for how a maintainer swaps it in."""
from tclip_b200.methods.dirichlet import BASE, EM_DIRICHLET  # noqa: F401
