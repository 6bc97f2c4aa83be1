import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
import torch
from yogo_b200 import _lib as L
lib = L.lib(); dev = "cuda:0"
for (N,H,W,Cin,Cout,s) in [(1,8,32,16,32,1),(2,19,35,16,32,1),(1,8,64,32,64,2),(3,21,37,32,64,2),(64,386,516,16,32,1),(64,386,516,32,64,2)]:
    g = torch.Generator().manual_seed(1)
    Ho, Wo = (H-1)//s+1, (W-1)//s+1
    x = torch.randn(N,H,W,Cin,generator=g).to(dev).bfloat16()
    dz = torch.randn(N,Ho,Wo,Cout,generator=g).to(dev).bfloat16()
    res = {}
    for name,opt in (("tc", 24601 + 131072), ("hmma", 24601 + 65536)):
        lib.yg_set_tc_options(opt)
        dw = torch.zeros(Cout,Cin,3,3,device=dev); db = torch.zeros(Cout,device=dev)
        nb = lib.yg_conv_wgrad_workspace(N,H,W,Cin,Cout,3,s); ws = torch.empty(max(nb,16),dtype=torch.uint8,device=dev)
        L.check(lib.yg_conv_wgrad(x.data_ptr(),dz.data_ptr(),dw.data_ptr(),db.data_ptr(),1,N,H,W,Cin,Cout,3,s,0.0,ws.data_ptr(),nb,L.stream()))
        torch.cuda.synchronize()
        e0,e1 = torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            L.check(lib.yg_conv_wgrad(x.data_ptr(),dz.data_ptr(),dw.data_ptr(),db.data_ptr(),1,N,H,W,Cin,Cout,3,s,0.0,ws.data_ptr(),nb,L.stream()))
        e1.record(); torch.cuda.synchronize()
        res[name] = (dw.clone(), db.clone(), e0.elapsed_time(e1)/3)
    lib.yg_set_tc_options(24601)
    a,b = res["tc"],res["hmma"]
    print({"shape":(N,H,W,Cin,Cout,s),"dw_rel":float((a[0]-b[0]).norm()/a[0].norm()),"db_rel":float((a[1]-b[1]).norm()/a[1].norm()),"ms_tc":round(a[2],4),"ms_hmma":round(b[2],4)}, flush=True)
