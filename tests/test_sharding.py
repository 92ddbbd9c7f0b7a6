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
