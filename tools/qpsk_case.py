import sys; sys.path.insert(0,'.')
import tools.bench_configs as b, torch, numpy as np, json, ctypes as Ct
import qradiolink_b200 as q
dev=torch.device("cuda",0); rng=np.random.default_rng(7)
C,T=256,1<<20; nb=T//32
txq=q.make_gr_mod_qpsk(4,1000000,1700,160000,n_channels=C,max_items=nb)
dq=torch.from_numpy(rng.integers(0,256,(C,nb),dtype=np.uint8)).to(dev)
txq.work_device(dq.data_ptr(),nb,nb); txq.sync()
L=q.load_library(); nout=nb*32
Xt=torch.empty((C,nout),dtype=torch.complex64,device=dev); n2=Ct.c_long()
assert L.qrl_tx_read(txq._h,Ct.c_void_p(Xt.data_ptr()),nout,Ct.byref(n2),1)==0
X=torch.empty((C,T),dtype=torch.complex64,device=dev); X.copy_(Xt[:,:T]*0.5)
X+=torch.view_as_complex(torch.randn((C,T,2),device=dev)*0.03); del Xt; txq.close()
b.rx_case("256ch QPSK-250k RX (config 3), T=2^20 per call", q.make_gr_demod_qpsk,(2,1000000,1700,160000),C,T,X,12.0,k=3)
