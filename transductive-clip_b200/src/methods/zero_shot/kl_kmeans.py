"""Drop-in for the reference's ``src/methods/zero_shot/kl_kmeans.py`` (imported by ``src/eval_zero_shot.py:12-19``): same
class name, constructor and ``run_task(task_dic)``; the arithmetic runs in libtclip_b200 (sm_100a CUDA)."""
from tclip_b200.methods.kmeans import BASE, KL_KMEANS  # noqa: F401
