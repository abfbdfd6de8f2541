"""The launches profiled for profiles/r02 (ncu --set full): the one-kernel search at batch 1 on 10M x 768 bf16, the 2-SM MMA
sweep at batch 4096, and the vectorised ingest kernel from a device source.  Small and deterministic on purpose.
    python scripts/profile_targets.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:"sweep_fused_kernel|gemm_pair_kernel|ingest_vec_kernel" -c 7 \
        -o gpurun_out/prof_r02 python scripts/profile_targets.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragfin_b200
from ragfin_b200.synthetic import synth_rows
rows = 10_000_000
idx = ragfin_b200.Index(768, "bf16", capacity=rows + 1_000_000)
for r in range(0, rows, 1_000_000):
    idx.add_synthetic(1234, r, 1_000_000)
q1 = torch.from_numpy(synth_rows(1235, 0, 1, 768)).cuda()
q4096 = torch.from_numpy(synth_rows(1235, 0, 4096, 768)).cuda()
for _ in range(3):                      # launches 1-3: sweep_fused_kernel
    ids, sc = idx.search_device(q1, 10)
torch.cuda.synchronize()
print("batch 1:", idx.stats(), ids[0, :3].tolist())
for _ in range(2):                      # launches 4-5: gemm_pair_kernel (append mode)
    ids, sc = idx.search_device(q4096, 10)
torch.cuda.synchronize()
print("batch 4096:", idx.stats(), ids[0, :3].tolist())
src = torch.randn((1_000_000, 768), dtype=torch.float32, device="cuda")
tmp = ragfin_b200.Index(768, "bf16", capacity=2_000_000)
for _ in range(2):                      # launches 6-7: ingest_vec_kernel
    tmp.add(src)
torch.cuda.synchronize()
print("ingest ok", len(tmp))
