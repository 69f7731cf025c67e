#!/usr/bin/env python3
"""Where does a tile's time go in the live int8-image scan?  Runs a few searches on a PKV_TILE_TIMING build
(PKV_LIB_PATH=build_dbg/libpkv_timing.so) and prints, per warp role, the average clocks per tile spent waiting for the
accumulator, loading it, and processing it."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import panoptikon_b200 as pk  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
ix = pk.VectorIndex(768, pk.F32, device=0)
ix.reserve(rows)
for b, off, n in bench.corpus_blocks(0, rows):
    ix.append(bench.gen_block(torch, b, bench.BLOCK_ROWS, 768, dev)[off:off + n].contiguous())
ix.seal()
q = bench.gen_queries(torch, batch, 768, dev)
for _ in range(4):
    ix.search(q, 100, pk.COSINE)
torch.cuda.synchronize()
L = pk.lib()
out = np.zeros(256 * 20 * 14, dtype=np.uint64)
L.pkv_debug_tile_timing.argtypes = [C.c_void_p, C.c_size_t]
rc = L.pkv_debug_tile_timing(out.ctypes.data, out.size)
t = out[:256 * 20 * 8].reshape(256, 20, 8).astype(np.float64)
t2 = out[256 * 20 * 8:256 * 20 * 12].reshape(256, 20, 4).astype(np.float64)
t3 = out[256 * 20 * 12:].reshape(256, 20, 2).astype(np.float64)
mma = t[:, 1, :]
mma = mma[mma[:, 0] > 0]
print(f"rc {rc}; MMA warps {len(mma)}: tiles {mma[:,0].mean():.0f}  wait tmem_empty {np.mean(mma[:,1]/mma[:,0]):.0f}  "
      f"wait full {np.mean(mma[:,2]/mma[:,0]):.0f}  issue+waits {np.mean(mma[:,3]/mma[:,0]):.0f} clk/tile")
epi = t[:, 2:18, :]
ok = epi[:, :, 0] > 0
n = epi[:, :, 0][ok]
print(f"epilogue warps {ok.sum()}: tiles {n.mean():.0f}  wait tmem_full {np.mean(epi[:,:,1][ok]/n):.0f}  ld+arrive "
      f"{np.mean(epi[:,:,2][ok]/n):.0f}  process+bound {np.mean(epi[:,:,3][ok]/n):.0f} clk/tile; slowest single tile "
      f"{epi[:,:,4][ok].mean():.0f} (max {epi[:,:,4][ok].max():.0f}); tiles > 1500 clk: {100*np.mean(epi[:,:,5][ok]/n):.1f} %")
per_smsp = [np.mean((epi[:, i::4, 3][ok[:, i::4]]) / epi[:, i::4, 0][ok[:, i::4]]) for i in range(4)]
print("process+bound by warp%4 (scheduler):", [f"{x:.0f}" for x in per_smsp])
per_rank = [np.mean(epi[r::2, :, 1][ok[r::2]] / epi[r::2, :, 0][ok[r::2]]) for r in range(2)]
print("wait tmem_full by CTA rank:", [f"{x:.0f}" for x in per_rank])
fl = t2[:, 2:18, :]
print(f"loop top (thresholds, row figures, bound): {np.mean(epi[:,:,6][ok]/n):.0f} clk/tile, > 1500 clk in {100*np.mean(epi[:,:,7][ok]/n):.2f} % of tiles")
nf = fl[:, :, 1][ok]
if nf.sum() > 0:
  print(f"flush_held: {nf.mean():.1f} calls per warp, {np.sum(fl[:,:,0][ok])/max(nf.sum(),1):.0f} clk per call, slowest {fl[:,:,2][ok].max():.0f}; "
      f"{np.mean(fl[:,:,0][ok]/n):.0f} clk/tile")
f3 = t3[:, 2:18, :]
tot = max(nf.sum(), 1)
if nf.sum() > 0:
  print(f"  inside a flush: round (smem, ring) {np.sum(fl[:,:,3][ok])/tot:.0f}  park_complete {np.sum(f3[:,:,0][ok])/tot:.0f}  atomic issue {np.sum(f3[:,:,1][ok])/tot:.0f} clk")
ix.close()
