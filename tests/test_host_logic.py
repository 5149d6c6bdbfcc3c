"""CPU: host-side logic of the product package — synthetic task generator layout, batch sharding (gloo, world_size 2),
label matching vs the oracle's, the drop-in module names, and the refusal to run without a B200."""
from __future__ import annotations

import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from oracle import restated as R
from oracle.ref_loader import make_args
import scipy_matching as matching
from tclip_b200 import tasks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_zero_shot_batch_layout():
    td, txt = tasks.make_zero_shot_batch(4, 50, seed=3)
    assert td["x_q"].shape == (4, 75, 50) and td["x_q"].dtype == torch.float32 and td["x_q"].is_contiguous()
    assert td["y_q"].shape == (4, 75, 1) and td["y_q"].dtype == torch.int64
    assert torch.allclose(td["x_q"].sum(-1), torch.ones(4, 75), atol=1e-5)       # softmax features
    assert txt.shape == (50, 1024)
    for t in range(4):
        assert 3 <= td["y_q"][t].unique().numel() <= 10                           # k_eff ~ U{3..10}
    td2, _ = tasks.make_zero_shot_batch(4, 50, seed=3)
    assert torch.equal(td["x_q"], td2["x_q"]) and torch.equal(td["y_q"], td2["y_q"])
    td3, _ = tasks.make_zero_shot_batch(4, 50, seed=3, batch_index=1)
    assert not torch.equal(td["x_q"], td3["x_q"])


def test_few_shot_batch_layout():
    td, _ = tasks.make_few_shot_batch(2, 20, shots=4, seed=1)
    assert td["x_s"].shape == (2, 80, 20) and td["y_s"].shape == (2, 80, 1)
    for t in range(2):
        assert torch.equal(td["y_s"][t].flatten().bincount(minlength=20), torch.full((20,), 4))
        assert td["y_q"][t].unique().numel() <= 5


def test_shard_batches_partition():
    for n, w in ((13, 1), (13, 2), (1333, 8), (3, 8)):
        parts = [tasks.shard_batches(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_matching_equals_oracle_matching():
    g = torch.Generator().manual_seed(0)
    T, n, K = 5, 75, 40
    feats = torch.softmax(3 * torch.randn(T, n, K, generator=g), -1)
    labels = torch.randint(0, 9, (T, n), generator=g) * 4
    # what tclip_cluster_prototypes delivers: clusters in first-appearance order, their mean raw feature
    proto = np.zeros((T, n, K), dtype=np.float32)
    n_clusters = np.zeros(T, dtype=np.int32)
    sample_cluster = np.zeros((T, n), dtype=np.int32)
    for t in range(T):
        order = []
        for i in range(n):
            lab = int(labels[t, i])
            if lab not in order:
                order.append(lab)
            sample_cluster[t, i] = order.index(lab)
        n_clusters[t] = len(order)
        for c, lab in enumerate(order):
            proto[t, c] = feats[t][labels[t] == lab].mean(0).numpy()
    onehot = torch.nn.functional.one_hot(labels, K).float()
    _, protos = R.cluster_prototypes(onehot, feats, K, "einsum")
    assert (matching.graph_matching(proto, n_clusters, sample_cluster) == R.graph_matching(labels, protos, K).numpy()).all()
    assert (matching.basic_matching(proto, n_clusters, sample_cluster) == R.basic_matching(labels, protos).numpy()).all()


def test_drop_in_module_names():
    """The evaluators import these dotted names (src/eval_zero_shot.py:12-13, src/eval_few_shot.py:12-13)."""
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {os.path.join(ROOT, 'transductive-clip_b200')!r})
        from src.methods.zero_shot.em_dirichlet import EM_DIRICHLET as A, BASE
        from src.methods.zero_shot.hard_em_dirichlet import HARD_EM_DIRICHLET as B
        from src.methods.few_shot.em_dirichlet import EM_DIRICHLET as C
        from src.methods.few_shot.hard_em_dirichlet import HARD_EM_DIRICHLET as D
        from src.methods.zero_shot.soft_kmeans import SOFT_KMEANS as E
        from src.methods.zero_shot.hard_kmeans import HARD_KMEANS as F
        from src.methods.zero_shot.em_gaussian import EM_GAUSSIAN as G
        from src.methods.zero_shot.em_gaussian_cov import EM_GAUSSIAN_COV as H
        from src.methods.zero_shot.kl_kmeans import KL_KMEANS as I
        import inspect
        for cls in (A, B, C, D, E, F, G, H, I):
            assert list(inspect.signature(cls.__init__).parameters)[1:] == ['model', 'device', 'log_file', 'args']
        assert list(inspect.signature(A.run_task).parameters)[1:] == ['task_dic']
        assert list(inspect.signature(C.run_task).parameters)[1:] == ['task_dic', 'shot']
        print('ok')
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_no_cpu_fallback():
    from tclip_b200 import ops
    from tclip_b200.methods.dirichlet import EM_DIRICHLET, FEW_SHOT_EM_DIRICHLET
    m = EM_DIRICHLET(model=None, device=torch.device("cpu"), log_file=None, args=make_args(20, iters=2))
    assert m.lambd == int(20 / 5) * 75
    td, _ = tasks.make_zero_shot_batch(2, 20, seed=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.run_task(td)
    f = FEW_SHOT_EM_DIRICHLET(model=None, device=torch.device("cpu"), log_file=None, args=make_args(20, iters=2, k_eff=4))
    assert f.lambd == int(20 / 4) * 75
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.log_features(torch.ones(4))
    with pytest.raises(ValueError, match="mm_mode"):
        EM_DIRICHLET(model=None, device=torch.device("cpu"), log_file=None, args=make_args(20, mm_mode="bogus"))
    from tclip_b200.pipeline import BatchPipeline
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchPipeline(torch.device("cpu"), streams=2)
    for bad in ("contraction", "match_clusters"):
        assert hasattr(ops, bad)
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.contraction(torch.zeros(1, 3, 4), torch.ones(1, 4, 4))
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.match_clusters(torch.zeros(1, 3, 4), torch.zeros(1, dtype=torch.int32), torch.zeros(1, 3, dtype=torch.int32))


def test_zero_shot_sampler_mirrors_the_reference():
    """tasks.ZeroShotQuerySampler draws the same index lists as the reference's CategoriesSampler_zero_shot +
    SamplerQuery_zero_shot under the same seeds (needs the reference checkout: build container only)."""
    import random
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout not present")
    mod = ref_loader.load("sampler_zero_shot")
    g = torch.Generator().manual_seed(3)
    n_class, n_query, n_batch = 40, 75, 6
    labels = torch.randint(0, n_class, (5000,), generator=g)
    def draw(make):
        random.seed(11)
        torch.manual_seed(12)
        return [q.clone() for q in make()]
    def ref():
        cs = mod.CategoriesSampler_zero_shot(n_batch, 5, n_class, n_query, force_query_size=True)
        cs.create_list_classes(labels)
        return mod.SamplerQuery_zero_shot(cs)
    want = draw(ref)
    got = draw(lambda: tasks.ZeroShotQuerySampler(n_batch, n_class, n_query, labels, force_query_size=True))
    assert len(want) == len(got) == n_batch
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tasks.DeviceTaskSource(torch.zeros(4, 3), torch.zeros(4), torch.device("cpu"))


def test_few_shot_samplers_mirror_the_reference():
    import random
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout not present")
    mod = ref_loader.load("sampler_few_shot")
    g = torch.Generator().manual_seed(4)
    n_class, n_query, n_batch, shots, k_eff = 12, 75, 4, 3, 5
    ls = torch.cat([torch.arange(n_class).repeat(6), torch.randint(0, n_class, (200,), generator=g)])
    lq = torch.randint(0, n_class, (4000,), generator=g)
    def seeded():
        random.seed(5)
        torch.manual_seed(6)
    seeded()
    cs = mod.CategoriesSampler_few_shot(n_batch, k_eff, n_class, shots, n_query, force_query_size=True)
    cs.create_list_classes(ls, lq)
    want_q = [q.clone() for q in mod.SamplerQuery_few_shot(cs)]
    want_s = [q.clone() for q in mod.SamplerSupport_few_shot(cs)]
    seeded()
    mine = tasks.FewShotSamplers(n_batch, k_eff, n_class, shots, n_query, ls, lq)
    got_q = [q.clone() for q in mine.query()]
    got_s = [q.clone() for q in mine.support()]
    assert all(torch.equal(a, b) for a, b in zip(want_q, got_q)) and len(got_q) == n_batch
    assert all(torch.equal(a, b) for a, b in zip(want_s, got_s)) and len(got_s) == n_batch


def test_few_shot_host_construction_equals_reference_generator():
    from oracle import ref_loader
    from host_twin import few_shot_tasks_on_host
    if not ref_loader.available():
        pytest.skip("reference checkout not present")
    g = torch.Generator().manual_seed(22)
    n_class, n_query, T, shots, k_eff = 16, 75, 3, 2, 5
    fs = torch.softmax(2 * torch.randn(400, n_class, generator=g), -1)
    ls = torch.cat([torch.arange(n_class).repeat(5), torch.randint(0, n_class, (320,), generator=g)])
    fq = torch.softmax(2 * torch.randn(3000, n_class, generator=g), -1)
    lq = torch.randint(0, n_class, (3000,), generator=g)
    torch.manual_seed(9)
    smp = tasks.FewShotSamplers(T, k_eff, n_class, shots, n_query, ls, lq)
    idx_q, idx_s = list(smp.query()), list(smp.support())
    want = few_shot_tasks_on_host(fs, ls, fq, lq, idx_s, idx_q)
    gen = ref_loader.load("task_generator_few_shot").Tasks_Generator_few_shot(
        k_eff=k_eff, shot=shots, n_query=n_query, n_class=n_class,
        loader_support=[(fs[i, :], ls[i]) for i in idx_s], loader_query=[(fq[i, :], lq[i]) for i in idx_q],
        model=None, args=make_args(n_class, use_softmax_feature=True))
    ref = gen.generate_tasks()
    for k in want:
        assert torch.equal(ref[k].reshape(want[k].shape), want[k]), k


def test_feature_cache_format_roundtrip(tmp_path):
    """tclip_b200.features writes / reads the reference's .plk cache (src/utils.py:241-249,298-306): our file loads with the
    reference's load_pickle and vice versa; paths follow the reference's naming."""
    from tclip_b200 import features
    g = torch.Generator().manual_seed(1)
    feats = torch.softmax(torch.randn(50, 7, generator=g), -1)
    labels = torch.randint(0, 7, (50,), generator=g)
    p = features.softmax_features_path("caltech101", "test", "RN50", 30, root=str(tmp_path))
    assert p.endswith(os.path.join("caltech101", "saved_features", "test_softmax_RN50_T30.plk"))
    assert features.visual_features_path("caltech101", "val", "RN50", root="data") == \
        os.path.join("data", "caltech101", "saved_features", "val_visual_RN50.plk")
    features.save_features(p, feats, labels)
    f2, l2 = features.load_features(p)
    assert torch.equal(f2, feats) and torch.equal(l2, labels) and l2.dtype == torch.int64
    from oracle import ref_loader
    if ref_loader.available():
        utils = ref_loader.load("utils")
        d = utils.load_pickle(p)                                  # the reference reads our file
        assert torch.equal(d["concat_features"], feats) and torch.equal(d["concat_labels"], labels)
        q = str(tmp_path / "ref.plk")
        utils.save_pickle(q, {"concat_features": feats, "concat_labels": labels.int()})
        f3, l3 = features.load_features(q)                        # we read the reference's file
        assert torch.equal(f3, feats) and torch.equal(l3, labels)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_gather_gloo_world2(tmp_path):
    """The N > 1 host path of bench.py on CPU: batches by rank, one gather of accuracies, max-over-ranks timing."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {os.path.join(ROOT, 'transductive-clip_b200')!r})
        import torch, torch.distributed as dist
        from tclip_b200 import tasks
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        mine = tasks.shard_batches(7, rank, world)
        acc = torch.tensor([float(sum(mine))])                    # stands in for the per-rank accuracy
        gathered = [torch.zeros(1) for _ in range(world)] if rank == 0 else None
        dist.gather(acc, gathered, dst=0)
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert sorted(sum([tasks.shard_batches(7, r, world) for r in range(world)], [])) == list(range(7))
            assert float(torch.cat(gathered).sum()) == 21.0 and float(t) == float(world)
            print("gather ok")
        dist.destroy_process_group()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "gather ok" in out.stdout, out.stderr[-2000:]


def test_bench_reference_arm_and_no_gpu_behaviour():
    """bench.py: the reference arm (CPU port on the host cores) prints one JSON line with the contract's keys; the B200 arm
    refuses to run without a CUDA device instead of falling back."""
    import json
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--classes", "30",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "tasks/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         env=env, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


# ----------------------------------------------------------------------------------------------------------------------
# evaluator-level boundary: the restated call sequence == the live evaluator, and the live evaluator
# reaches the drop-in classes through its own get_method_builder
# ----------------------------------------------------------------------------------------------------------------------
class _RecordingMethod:
    """Stands in for a method class: keeps what the evaluator hands to run_task."""
    seen = None

    def __init__(self, model, device, log_file, args):
        self.args = args

    def run_task(self, task_dic):
        type(self).seen.append({k: v.clone() for k, v in task_dic.items()})
        n_task = task_dic["x_q"].shape[0]
        acc = (task_dic["x_q"].argmax(-1) == task_dic["y_q"].squeeze(2)).float().mean(1, keepdim=True).numpy()
        return {"acc": acc, "timestamps": 0.5 * n_task, "criterions": np.zeros(1)}


def _evaluator_inputs():
    K, per_class = 30, 30      # >= 25 per class: three classes must hold n_query samples or the sampler re-draws forever
    g = torch.Generator().manual_seed(0)
    labels = torch.arange(K).repeat_interleave(per_class)
    feats = torch.softmax(3.0 * torch.randn(labels.numel(), K, generator=g) + 4.0 * torch.nn.functional.one_hot(labels, K), -1)
    args = make_args(K, iters=2, name_method="EM_DIRICHLET", number_tasks=20, batch_size=5, used_test_set="test",
                     dataset="synthetic", save_results=False)
    return feats, labels, args


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["available"]).available(), reason="reference checkout not present")
def test_evaluator_harness_equals_live_evaluator(tmp_path):
    """The reference's own Evaluator_zero_shot.evaluate_tasks (imported unchanged) and tests/evaluator_harness.py hand the
    same task dictionaries to run_task and return the same means, given the same seeds."""
    import random

    import evaluator_harness as H
    from oracle import ref_loader
    ev_mod = ref_loader.load("eval_zero_shot")
    feats, labels, args = _evaluator_inputs()
    live = type("Live", (_RecordingMethod,), {"seen": []})
    mine = type("Mine", (_RecordingMethod,), {"seen": []})
    ev = ev_mod.Evaluator_zero_shot(torch.device("cpu"), args, str(tmp_path / "log.txt"))
    ev.get_method_builder = lambda model, device, args, log_file: live(model, device, log_file, args)
    random.seed(11), torch.manual_seed(11)
    acc_live, t_live = ev.evaluate_tasks(None, feats, labels)
    random.seed(11), torch.manual_seed(11)
    acc_mine, t_mine, _ = H.evaluate_tasks(args, torch.device("cpu"), feats, labels, build=lambda **kw: mine(**kw))
    assert len(live.seen) == len(mine.seen) == 4
    for a, b in zip(live.seen, mine.seen):
        assert torch.equal(a["x_q"], b["x_q"]) and torch.equal(a["y_q"], b["y_q"])
        assert a["x_q"].shape == (5, 75, 30) and a["y_q"].shape == (5, 75, 1) and a["y_q"].dtype == torch.int64
    assert acc_live == acc_mine and t_live == t_mine


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["available"]).available(), reason="reference checkout not present")
def test_live_evaluator_reaches_the_drop_in_class(tmp_path):
    """The unchanged evaluator, with the drop-in class bound where its import statement would have bound the reference's
    (src/eval_zero_shot.py:14): its own get_method_builder constructs the B200 class with its own keyword call and calls
    run_task on the task_dic of its own task generator; without a GPU that call must refuse loudly (no CPU path)."""
    from oracle import ref_loader
    from tclip_b200.methods.dirichlet import EM_DIRICHLET
    ev_mod = ref_loader.load("eval_zero_shot")
    feats, labels, args = _evaluator_inputs()
    ev = ev_mod.Evaluator_zero_shot(torch.device("cpu"), args, str(tmp_path / "log.txt"))
    original = ev_mod.EM_DIRICHLET
    ev_mod.EM_DIRICHLET = EM_DIRICHLET
    try:
        with pytest.raises(RuntimeError, match="B200 GPUs only"):
            ev.evaluate_tasks(None, feats, labels)
    finally:
        ev_mod.EM_DIRICHLET = original


def test_logger_handlers_are_shared_and_released_without_blocking(tmp_path):
    """One handler per (logger name, sink) however many method objects are alive (a new one is built per batch, several at
    once under tclip_b200.pipeline); a sink closes with its last user; a release arriving while the registry lock is held
    (the garbage collector running __del__ inside the critical section) is queued instead of deadlocking."""
    import logging

    from tclip_b200 import logger as L
    name, path = "tclip_test_logger", str(tmp_path / "log.txt")
    a, b = L.Logger(name, path), L.Logger(name, path)
    py = logging.getLogger(name)
    assert len(py.handlers) == 2                         # console + file, not four
    a.info("one line")
    a.del_logger()
    assert len(py.handlers) == 2                         # b still uses both
    assert L._LOCK.acquire(blocking=False)               # as if __del__ ran inside _acquire's critical section ...
    try:
        b.del_logger()                                   # ... must return at once
        assert len(py.handlers) == 2 and len(L._PENDING) == 2
    finally:
        L._LOCK.release()
    c = L.Logger(name, None)                             # the next user of the registry applies the queued releases
    assert [type(h) for h in py.handlers] == [logging.StreamHandler]
    c.del_logger()
    assert py.handlers == []
    assert open(path).read().count("one line") == 1


def test_bench_kmeans_loop_flop_count():
    """bench.py's FP32 work model of the k-means loop in sample coordinates (roofline.loop_fp32): dense 5 K r n flop per task and
    iteration; the triangular form counts the 16-blocks on and below the diagonal (+ the running sum of squares); features
    used as coordinates (D <= n) have no triangular form."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    r = bench.km_loop_fp32(75, 1000, 1024, 100, 20, 2.0, 73.0)
    assert r["dense_flop_per_task_iteration"] == 5.0 * 1000 * 75 * 75
    blocks = [(min(16, 75 - 16 * b), b) for b in range(5)]
    tri_c = sum(sz * 16 * (b + 1) for sz, b in blocks)
    tri_d = sum(sz * 16 * (5 - b) for sz, b in blocks)
    assert r["flop_per_task_iteration"] == 2.0 * 1000 * tri_c + 3.0 * 1000 * tri_d + 2.0 * 1000 * 75
    assert 0.6 < r["flop_per_task_iteration"] / r["dense_flop_per_task_iteration"] < 0.7
    np.testing.assert_allclose(r["achieved"], r["flop_per_task_iteration"] * 100 * 20 / 2.0e-3 / 1e12)
    np.testing.assert_allclose(r["frac"], r["achieved"] / 73.0)
    d = bench.km_loop_fp32(75, 30, 24, 10, 5, 1.0, 73.0)          # D <= n: r = D, dense
    assert d["flop_per_task_iteration"] == d["dense_flop_per_task_iteration"] == 5.0 * 30 * 24 * 75


def test_chained_softmax_algebra_float32():
    """The algebra of the chained k-means launches (csrc/kmeans_run.cu, ChainArgs) restated in numpy float32: per class tile
    of 128 the maximum m, e = exp(l - m) and s = sum e; then M = max m, S = sum_tiles s exp(m - M) and u = e exp(m - M) / S.
    Against the float64 soft-max of the same float32 logits the result is within a few ulp of the largest probability — the
    same bound as the one-pass float32 soft-max — also when a tile lies far below the maximum (its factor underflows to 0)."""
    rng = np.random.default_rng(0)
    for K, spread in ((1000, 30.0), (131, 5.0), (600, 200.0)):
        l = (rng.standard_normal((75, K)) * spread).astype(np.float32)
        tiles = [(a, min(K, a + 128)) for a in range(0, K, 128)]
        m = np.stack([l[:, a:b].max(1) for a, b in tiles], 1)
        e = [np.exp(l[:, a:b] - m[:, [i]], dtype=np.float32) for i, (a, b) in enumerate(tiles)]
        s = np.stack([x.sum(1, dtype=np.float32) for x in e], 1)
        M = m.max(1, keepdims=True)
        S = np.zeros(75, np.float32)
        for i in range(len(tiles)):                       # tile order, like the kernel
            S = S + s[:, i] * np.exp(m[:, i] - M[:, 0], dtype=np.float32)
        u = np.concatenate([e[i] * (np.exp(m[:, [i]] - M, dtype=np.float32) / S[:, None]).astype(np.float32)
                            for i in range(len(tiles))], 1)
        l64 = l.astype(np.float64)
        ref = np.exp(l64 - l64.max(1, keepdims=True))
        ref /= ref.sum(1, keepdims=True)
        one_pass = np.exp(l - l.max(1, keepdims=True), dtype=np.float32)
        one_pass = one_pass / one_pass.sum(1, dtype=np.float32, keepdims=True)
        err_chain = np.abs(u - ref).max()
        err_one = np.abs(one_pass - ref).max()
        assert err_chain <= 8 * np.finfo(np.float32).eps, (K, spread, err_chain)
        assert err_chain <= max(4 * err_one, 4 * np.finfo(np.float32).eps), (K, spread, err_chain, err_one)
        assert (u.argmax(1) == ref.argmax(1)).all()
        np.testing.assert_allclose(u.sum(1), 1.0, atol=1e-5)


def test_sample_coordinate_identities_float64():
    """The two identities the k-means loop of csrc/kmeans_run.cu rests on, in numpy float64:
    (1) with G = X X^T = L L^T, a centroid w = c^T X (a combination of the task's samples) and wt = c^T L satisfy
        ||w - x_n||^2 = ||wt - L_n||^2 for every sample n — also when samples are duplicated (singular G, zero pivots);
    (2) L is lower triangular, so the distance of sample n splits at any block boundary b > n into the direct differences
        over the coordinates below b and the plain sum of wt_j^2 over the coordinates from b on (the triangular form)."""
    rng = np.random.default_rng(1)
    n, D, K = 40, 96, 7
    X = rng.standard_normal((n, D))
    X[5] = X[2]                                   # duplicated sample
    X[9] = 0.5 * (X[0] + X[3])                    # dependent sample
    G = X @ X.T
    # Cholesky with zeroed columns at (numerically) zero pivots, as chol_kernel does
    A = G.copy()
    diag = np.diag(G).copy()
    L = np.zeros_like(G)
    for j in range(n):
        p = A[j, j]
        d = np.sqrt(p) if (p > 1e-9 * diag[j] and p > 0) else 0.0
        if d > 0:
            L[j:, j] = A[j:, j] / d
            L[j, j] = d
            A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    np.testing.assert_allclose(L @ L.T, G, atol=1e-8)
    c = rng.random((n, K))
    c /= c.sum(0, keepdims=True)
    W, Wt = c.T @ X, c.T @ L
    d_feat = ((X[:, None, :] - W[None, :, :]) ** 2).sum(-1)
    d_coord = ((L[:, None, :] - Wt[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(d_coord, d_feat, rtol=1e-7, atol=1e-7)
    assert np.allclose(np.triu(L, 1), 0.0)
    for b in (16, 32):
        rows = np.arange(n) < b                   # samples of the groups below the boundary have no coordinate >= b
        split = ((L[rows][:, None, :b] - Wt[None, :, :b]) ** 2).sum(-1) + (Wt[:, b:] ** 2).sum(-1)[None, :]
        np.testing.assert_allclose(split, d_coord[rows], rtol=1e-12, atol=1e-12)


def test_diagnostic_switches_are_documented():
    """Every TCLIP_* environment variable the library or the Python package reads appears in INTEGRATION.md §9."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "transductive-clip_b200")
    names = set()
    for path in glob.glob(os.path.join(pkg, "csrc", "*.cu")) + glob.glob(os.path.join(pkg, "csrc", "*.cuh")):
        names |= set(re.findall(r'getenv\("(TCLIP_[A-Z0-9_]+)"', open(path).read()))
    for path in glob.glob(os.path.join(pkg, "tclip_b200", "**", "*.py"), recursive=True):
        names |= set(re.findall(r'environ(?:\.get)?[\[(]\s*"(TCLIP_[A-Z0-9_]+)"', open(path).read()))
    assert {"TCLIP_KM_CHAIN", "TCLIP_KM_TRI", "TCLIP_SPARSE_SOFTMAX", "TCLIP_MM_MODE", "TCLIP_LIB"} <= names
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing


def test_every_entry_point_is_in_the_integration_table():
    """Every function include/tclip_b200.h declares is named in INTEGRATION.md §3 (`X_workspace_bytes` may ride on `X`)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "tclip_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(tclip_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 30
    missing = [n for n in names if n not in doc and not (n.endswith("_workspace_bytes") and
                                                         (n[:-len("_workspace_bytes")] in doc or n.replace("_workspace_bytes", "_run") in doc))]
    assert not missing, missing
