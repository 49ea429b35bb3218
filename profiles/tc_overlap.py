"""Does a TMEM-reading phase overlap MMAs that are already queued?  (tmla_tc_overlap_probe, one CTA, clock64)"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from three_mlagents_b200 import native as nat
out = torch.zeros(8, dtype=torch.int64, device="cuda")
print("   N  #mma  #ld | mma alone  ld alone | mma+ld   ld under mma | issue-lane cycles inside tcgen05.mma")
for n, n_mma, n_ld in ((256, 16, 16), (256, 16, 64), (256, 32, 64), (128, 32, 64), (16, 16, 16), (16, 64, 64)):
    for _ in range(2):
        nat.check(nat.lib.tmla_tc_overlap_probe(nat.ptr(out), n, n_mma, n_ld, nat.current_stream()))
        torch.cuda.synchronize()
    o = out.cpu().tolist()
    print(f"{n:4d} {n_mma:5d} {n_ld:4d} | {o[0]:9d} {o[1]:9d} | {o[2]:7d} {o[3]:12d} | {o[4]}")
