"""Write tests/golden/ckpt_ref_msgm_d2.pt with the UNMODIFIED reference's ``NN.save_checkpoint`` (NN.py:10-19).

Build container only (needs /root/reference):   python tests/golden/make_checkpoint_golden.py
A d=2 dense multiplicative SDE + MLP(NormalizeLogRadius) trained for two Adam iterations on the reference's own
``ssm`` loss, then saved by the reference.  tests/test_checkpoint.py loads the file with this package's
``load_checkpoint`` (and writes / reloads it with the package's ``save_checkpoint``).
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_live  # noqa: E402

ref = ref_live.load()
torch.manual_seed(0)
np.random.seed(0)
random.seed(0)
x_init = torch.randn(256, 2) * 1.5
base, gen, net = ref_live.build(ref, "msgm_dense", 2, x_init, "NormalizeLogRadius")
opt = torch.optim.Adam(gen.parameters(), lr=1e-3)
for it in range(2):
    opt.zero_grad()
    gen.ssm(torch.randn(16, 2)).mean().backward()
    opt.step()
path = os.path.join(HERE, "ckpt_ref_msgm_d2.pt")
ref.NN.save_checkpoint(path, gen, opt, 1)
# side file: what the reference object held that its checkpoint does not (G, L_G, r_T) + a probe evaluation
torch.manual_seed(5)
y, s = torch.randn(8, 2), torch.rand(8)
with torch.no_grad():
    a = net(y, s)
np.savez_compressed(os.path.join(HERE, "ckpt_ref_msgm_d2_side.npz"), G=base.G.numpy(), L_G=base.L_G.numpy(),
                    r_T=base.r_T.numpy(), y=y.numpy(), s=s.numpy(), a=a.numpy())
print("wrote", path, os.path.getsize(path), "bytes")
