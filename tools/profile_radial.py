"""Radial hidden layer, tensor-core vs packed-FP32 kernel at the 2AA bench size (for ncu / timing): python tools/profile_radial.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jamun_b200 import ops
E=282000; cap=E; L=6
dev="cuda"
rb=torch.rand(cap,32,device=dev); w=torch.randn(L,32,64,device=dev)/5.6; b=torch.randn(L,2,64,device=dev)*0.3
flag=(torch.rand(cap,device=dev)<0.2).to(torch.uint8); rowptr=torch.tensor([0,E],dtype=torch.int32,device=dev)
h=torch.empty(L,cap,64,device=dev); img=ops.radial_pack_frag(w)
for _ in range(3):
    ops.edge_radial_hidden_mma(rb,flag,rowptr,img,b,h); ops.edge_radial_hidden_all(rb,flag,rowptr,w,b,h)
torch.cuda.synchronize()
