"""How much of a lock-step round of agg_wh_quad_kernel is real work: the 32 stream items of a CTA item run
max(edges) / 4 rounds, every quad draws bytes for 4 tile rows per round whether its item still has edges or not."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import stag_b200 as sb

src, dst = bench.synth_graph()
g = sb.Graph(torch.from_numpy(src), torch.from_numpy(dst), bench.N_NODES).to("cuda")
for by_dst in (True, False):
    sg, keep = g._s.csx(by_dst)
    items = keep["items"].cpu().numpy() if isinstance(keep, dict) else None
    if items is None:
        print("no items tensor"); break
    hub_ptr = keep["hub_seg_ptr"].cpu().numpy(); hub_rows = keep["hub_rows"].cpu().numpy(); indptr = keep["indptr"].cpu().numpy()
    sizes = []
    for h, row in enumerate(hub_rows):
        d = int(indptr[row + 1] - indptr[row]); nseg = int(hub_ptr[h + 1] - hub_ptr[h])
        sizes += [min(128, d - 128 * k) for k in range(nseg)]
    e0, e1 = items[:, 2], items[:, 3]
    sz = np.where(e1 >= 0, e1, e0) - e0
    sizes = np.array(sizes + list(sz), dtype=np.int64)
    n = len(sizes); pad = (-n) % 32
    s32 = np.concatenate([sizes, np.zeros(pad, np.int64)]).reshape(-1, 32)
    rounds = (s32.max(1) + 3) // 4
    print("by_dst=%s: %d items (%d hub segments), edges %d, item size mean %.1f max %d; CTA items %d, rounds mean %.1f; "
          "tile rows drawn %d = %.3f x edges" % (by_dst, n, len(hub_rows) and int(hub_ptr[-1]), sizes.sum(), sizes.mean(), sizes.max(),
                                               len(s32), rounds.mean(), rounds.sum() * 128, rounds.sum() * 128 / sizes.sum()))
    q = np.percentile(rounds, [5, 25, 50, 75, 95, 100]); print("   rounds percentiles 5/25/50/75/95/100:", q)
    # schedule of the 16-sample launch over 2 x 148 resident CTAs: static stride (the kernel's loop) against a dynamic queue
    import heapq
    S, ncta = 16, 2 * 148
    work = np.tile(rounds + 1.2, S)      # + ~1.2 rounds of prologue per CTA item (bytes, MMA and gather latency exposed)
    per = np.zeros(ncta); np.add.at(per, np.arange(len(work)) % ncta, work)
    h = [0.0] * ncta; heapq.heapify(h)
    for w in work: heapq.heappush(h, heapq.heappop(h) + w)
    print("   static stride: max / mean CTA load %.3f;  dynamic queue in item order: %.3f" % (per.max() / per.mean(), max(h) / (work.sum() / ncta)))
