"""CPU tests of the N>1 host path: world_size-2 gloo processes shard a batch with
sa_partition_batch / sharding.py, align their shard (the oracle stands in for the GPU here --
test infrastructure only) and gather; the result must equal the single-process one."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from gpu_common import ROOT, load_package


def test_shard_csr_rebases_offsets():
    sa = load_package()
    if not os.path.exists(sa.LIB_PATH):
        sa.build()
    from sa_b200 import sharding
    rng = np.random.default_rng(1)
    n = rng.integers(5, 60, 40); m = rng.integers(5, 60, 40)
    toff = np.concatenate(([0], np.cumsum(n))); poff = np.concatenate(([0], np.cumsum(m)))
    T = rng.integers(0, 4, toff[-1], dtype=np.uint8); P = rng.integers(0, 4, poff[-1], dtype=np.uint8)
    part = sa.partition_batch(toff, poff, 3)
    seen = 0
    for r in range(3):
        sh = sharding.shard_csr(part, T, toff, P, poff, r)
        assert sh["first"] == seen and sh["text_off"][0] == 0 and sh["pattern_off"][0] == 0
        for k in range(sh["count"]):
            g = sh["first"] + k
            assert np.array_equal(sh["text"][sh["text_off"][k]:sh["text_off"][k + 1]], T[toff[g]:toff[g + 1]])
            assert np.array_equal(sh["pattern"][sh["pattern_off"][k]:sh["pattern_off"][k + 1]], P[poff[g]:poff[g + 1]])
        seen += sh["count"]
    assert seen == 40


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    sys.path.insert(0, os.path.join({root!r}, "sequence-alignment-gpu_b200"))
    from gpu_common import load_package
    sa = load_package()
    from sa_b200 import sharding
    from oracle.oracle_py import Oracle
    import synth
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, toff, P, poff = synth.synthetic_batch(120, seed=5, lo=30, hi=90)
    mat = np.full((23, 23), -2, np.int32); np.fill_diagonal(mat, 6)
    orc = Oracle()
    def align_fn(t, to, p, po):
        return [orc.align(1, 23, mat, 5, t[to[i]:to[i+1]], p[po[i]:po[i+1]]).key() for i in range(len(to) - 1)]
    part = sa.partition_batch(toff, poff, world)
    first, merged = sharding.align_batch_sharded(align_fn, part, T, toff, P, poff, rank, world, gather=True)
    if rank == 0:
        full = align_fn(T, toff, P, poff)
        assert merged == full, "sharded result differs from the single-process result"
        counts = np.diff(part.astype(np.int64)).tolist()
        print(json.dumps(dict(ok=True, counts=counts)))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_sharded_batch(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert '"ok": true' in outs[0][0]


def test_slice_columns_cover_the_text():
    load_package()
    from sa_b200 import strips
    for n, w in ((10, 1), (10, 3), (7, 8), (100001, 4), (952381, 8)):
        sl = strips.slice_columns(n, w)
        assert len(sl) == w and sl[0][0] == 0 and sum(x[1] for x in sl) == n
        for (a, wa), (b, _) in zip(sl, sl[1:]):
            assert a + wa == b


def test_strips_local_matches_oracle(oracle):
    """Host logic of the column-slice path (strips.py) with the numpy stand-in engine: 1, 2, 3 and 5 slices."""
    load_package()
    from sa_b200 import strips
    from helpers import NumpyStripEngine
    rng = np.random.default_rng(11)
    mat = np.full((4, 4), -4, np.int32); np.fill_diagonal(mat, 5)
    for n, m in ((57, 49), (40, 66), (5, 3)):
        t = rng.integers(0, 4, n, dtype=np.uint8); p = t[:m].copy() if m <= n else rng.integers(0, 4, m, dtype=np.uint8)
        p[::7] = (p[::7] + 1) % 4
        want = oracle.align(0, 4, mat, 5, t, p)
        for world in (1, 2, 3, 5):
            eng = [NumpyStripEngine(mat, 5, t[c0:c0 + w], c0, p, b"ATCG-") for c0, w in strips.slice_columns(n, world)]
            score, at, ap, ti, pi = strips.align_pair_strips_local(eng, m)
            assert (score, len(at), ti, pi, at, ap) == want.key(), (n, m, world)
            eng = [NumpyStripEngine(mat, 5, t[c0:c0 + w], c0, p, b"ATCG-") for c0, w in strips.slice_columns(n, world)]
            score, at, ap, ti, pi = strips.align_pair_strips_local(eng, m, chunks=3)
            assert (score, len(at), ti, pi, at, ap) == want.key(), (n, m, world, "row chunks")


STRIP_WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from gpu_common import load_package
    load_package()
    from sa_b200 import strips
    from helpers import NumpyStripEngine
    from oracle.oracle_py import Oracle
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    n, m = 83, 77
    t = rng.integers(0, 4, n, dtype=np.uint8); p = t[:m].copy(); p[::5] = (p[::5] + 2) % 4
    mat = np.full((4, 4), -4, np.int32); np.fill_diagonal(mat, 5)
    c0, w = strips.slice_columns(n, world)[rank]
    want = Oracle().align(0, 4, mat, 5, t, p)
    for chunks in (1, 4):
        eng = NumpyStripEngine(mat, 5, t[c0:c0 + w], c0, p, b"ATCG-"); eng.n_total = n
        got = strips.align_pair_strips(eng, m, rank, world, lambda: torch.zeros(m + 1, dtype=torch.int32), chunks=chunks)
        assert (got[0], len(got[1]), got[3], got[4], got[1], got[2]) == want.key(), "slices differ from the oracle"
    if rank == 0:
        print(json.dumps(dict(ok=True, score=got[0])))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_column_slices(tmp_path):
    script = tmp_path / "strip_worker.py"
    script.write_text(STRIP_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert '"ok": true' in outs[0][0]


def test_bench_spreads_ranks_over_both_host_bridges(monkeypatch):
    """bench.py maps local ranks to devices 0, D/2, 1, D/2+1, ... when a run uses fewer ranks than the box has GPUs
    (the end-to-end path is bound by the host side of PCIe, and the GPUs hang off two host bridges)."""
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    fake = types.SimpleNamespace(cuda=types.SimpleNamespace(device_count=lambda: 8))
    monkeypatch.setitem(sys.modules, "torch", fake)
    assert [bench.device_for_rank(r, 4) for r in range(4)] == [0, 4, 1, 5]
    assert [bench.device_for_rank(r, 2) for r in range(2)] == [0, 4]
    assert [bench.device_for_rank(r, 8) for r in range(8)] == list(range(8))
    monkeypatch.setenv("SA_BENCH_DEVICE_ORDER", "linear")
    assert [bench.device_for_rank(r, 4) for r in range(4)] == [0, 1, 2, 3]
    monkeypatch.delenv("SA_BENCH_DEVICE_ORDER")
    fake.cuda.device_count = lambda: 2
    assert [bench.device_for_rank(r, 2) for r in range(2)] == [0, 1]
