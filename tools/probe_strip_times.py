"""Per-strip timestamps of one long-pair fill (dev tool, needs a GPU): start, border ready, end of every strip, from the
kernel's globaltimer stamps (SA_LONG_DBG).   python tools/probe_strip_times.py [n] [m]    (SA_TILE / SA_LONG_R select the kernel)"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
import synth  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
import torch  # noqa: E402

sa = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t, p = synth.synthetic_pair(n, 12345, 54321)
if m:
    p = p[:m]
if os.environ.get("SLICE"):          # a column slice of the pair (config 5 shape: few columns, all rows)
    t = t[:int(os.environ["SLICE"])]
dbg = tempfile.mktemp(prefix="sa_dbg_")
os.environ["SA_LONG_DBG"] = dbg
mat = np.full((4, 4), -4, np.int32)
np.fill_diagonal(mat, 5)
al = sa.Aligner(0)
dT, dP = torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda()
right = torch.zeros(len(p) + 1, dtype=torch.int32, device="cuda")
for rep in range(3):
    al.strip_begin(4, mat, 5, dT.data_ptr(), len(t), 0, len(t), dP.data_ptr(), len(p), 0)
    al.strip_fill_rows(0, len(p), 0, right.data_ptr(), 0, 0)
    al.strip_linked_status()
raw = np.fromfile(dbg + ".dev0", dtype=np.uint64)
S = len(raw) // (16 + 128)
events = raw[16 * S:].reshape(S, 64, 2).astype(np.int64)      # the first 64 top-row stalls after the ramp: {first tile of the group, ns}
ts = raw[:3 * S].reshape(S, 3).astype(np.int64)
ex = raw[3 * S:7 * S].reshape(S, 4).astype(np.int64)          # loop start, k == 32, first drain group, last group
cnt = raw[7 * S:11 * S].reshape(S, 4).astype(np.int64)        # spin iterations, ns spent spinning, groups that stalled, of those in the ramp
ee = raw[11 * S:13 * S].reshape(S, 2).astype(np.int64)        # group of tiles 8001..8008: upkeep entered / left
wr = raw[13 * S:14 * S].astype(np.int64)                      # lane 31 stored tile 8008
t0 = ts[:, 0].min()
ts -= t0
ex -= t0
ee -= t0
wr -= t0
print(f"n={len(t)} m={len(p)} strips={S} SA_TILE={os.environ.get('SA_TILE')} SA_LONG_R={os.environ.get('SA_LONG_R')}")
print(f"kernel span {ts[:, 2].max() / 1e3:.1f} us; strip 0: start {ts[0, 0] / 1e3:.1f} ready {ts[0, 1] / 1e3:.1f} end {ts[0, 2] / 1e3:.1f} us")
ends = ts[:, 2] / 1e3
dur = (ts[:, 2] - ts[:, 1]) / 1e3
for s in sorted(set(list(range(min(S, 6))) + list(range(0, S, max(1, S // 12))) + [S - 1])):
    print(f"  strip {s:5d}: start {ts[s, 0] / 1e3:9.1f}  end {ends[s]:9.1f} us   end - end(prev) {ends[s] - ends[s - 1] if s else 0:7.2f} us")
for s in sorted(set(list(range(min(S, 4))) + [S // 2, S - 1])):
    print(f"  strip {s:5d}: loop start {ex[s, 0] / 1e3:9.1f}  k=32 {ex[s, 1] / 1e3:9.1f}  drain starts {ex[s, 2] / 1e3:9.1f}  last group {ex[s, 3] / 1e3:9.1f}  end {ends[s]:9.1f} us")
if S > 4 and ee[1, 0] > 0:
    for s in (1, 2, 3, S // 2):
        print(f"  strip {s}: producer stored tile 8008 at {wr[s - 1] / 1e3:.2f}, consumer entered the group's upkeep at {ee[s, 0] / 1e3:.2f}, left at {ee[s, 1] / 1e3:.2f} us"
              f"  (visible {(ee[s, 1] - wr[s - 1]) / 1e3:.2f} us after the store; own store of tile 8008 at {wr[s] / 1e3:.2f})")
if S > 1:
    print(f"  top-row waits per strip (median over strips >= 1): {np.median(cnt[1:, 2]):.0f} groups stalled ({np.median(cnt[1:, 3]):.0f} in the ramp), "
          f"{np.median(cnt[1:, 0]):.0f} reloads, {np.median(cnt[1:, 1]) / 1e3:.2f} us spinning")
    for name, col in (("loop start", ex[:, 0]), ("k=32", ex[:, 1]), ("drain start", ex[:, 2]), ("last group", ex[:, 3])):
        print(f"  lag per strip at {name}: median {np.median(np.diff(col)) / 1e3:.2f} us")
    d = np.diff(ends)
    print(f"end-to-end lag per strip: median {np.median(d):.2f} us, mean {d.mean():.2f} us, max {d.max():.2f} us")
for s_ in sorted(set([1, 2, 3, 4, 5, S // 2, S - 1])):
    if 0 < s_ < S:
        ev = events[s_]
        ev = ev[ev[:, 0] > 0]
        print(f"  strip {s_}: stalls after the ramp (group start tile : ns) " + " ".join(f"{a}:{b}" for a, b in ev[:24]))
        if len(ev) > 2:
            print(f"     gaps between stalled groups (tiles): {np.diff(ev[:, 0])[:24].tolist()}   median stall {np.median(ev[:, 1]):.0f} ns")
al.close()
