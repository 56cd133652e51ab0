"""Checkpoint format: a file written by the UNMODIFIED reference's ``save_checkpoint`` (fixture
tests/golden/ckpt_ref_msgm_d2.pt, made by tests/golden/make_checkpoint_golden.py) loads into this package's objects;
this package's files load back (round trip, including the SDE tensors the reference does not persist) and keep the
reference's six keys so that the reference's loader reads them.  Host logic only: runs without a GPU."""
import os

import numpy as np
import torch

import sdeflow_light_b200 as P
from sdeflow_light_b200 import NN

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_KEYS = {"iteration", "model", "optimizer", "torch_rng", "numpy_rng", "python_rng"}


def _objects(seed):
    torch.manual_seed(seed)
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    base = P.MSGMsde(torch.randn(256, 2) * 1.5, beta_min=0.1, beta_max=20., T=T, t_epsilon=1e-3, denseTensor=True,
                     norm_sampler="ecdf", norm_map="log", num_steps_forward=16, device="cpu",
                     estim_cst_norm_dens_r_T=False)
    gen = P.PluginReverseSDE(base, P.MLP(input_dim=2, index_dim=1, hidden_dim=128, premodule="NormalizeLogRadius"), T)
    return gen, torch.optim.Adam(gen.parameters(), lr=1e-3)


def test_reference_written_checkpoint_loads():
    gen, opt = _objects(123)
    it = NN.load_checkpoint(os.path.join(GOLD, "ckpt_ref_msgm_d2.pt"), gen, opt, "cpu")
    assert it == 1
    ck = torch.load(os.path.join(GOLD, "ckpt_ref_msgm_d2.pt"), map_location="cpu", weights_only=False)
    assert set(ck.keys()) == REF_KEYS
    for k, v in ck["model"].items():
        assert torch.equal(gen.state_dict()[k], v), k
    # Adam moments arrived: step counter 2 and non-zero exp_avg for every parameter
    st = opt.state_dict()["state"]
    assert len(st) == 8 and all(int(s["step"]) == 2 and float(s["exp_avg"].abs().sum()) > 0 for s in st.values())
    # the loaded weights evaluate like the reference's net did (torch-op forward on CPU tensors needs autograd mode on:
    # the no-grad path is the CUDA kernel)
    side = np.load(os.path.join(GOLD, "ckpt_ref_msgm_d2_side.npz"))
    y, s = torch.from_numpy(side["y"]), torch.from_numpy(side["s"])
    a = gen.a._forward_torch(y, s) if hasattr(gen.a, "_forward_torch") else None
    if a is not None:
        assert float((a.detach() - torch.from_numpy(side["a"])).abs().max()) < 1e-6


def test_round_trip_with_sde_tensors(tmp_path):
    gen, opt = _objects(7)
    NN.load_checkpoint(os.path.join(GOLD, "ckpt_ref_msgm_d2.pt"), gen, opt, "cpu")
    side = np.load(os.path.join(GOLD, "ckpt_ref_msgm_d2_side.npz"))
    gen.base_sde.G, gen.base_sde.L_G = torch.from_numpy(side["G"]), torch.from_numpy(side["L_G"])
    gen.base_sde.r_T = torch.from_numpy(side["r_T"])
    path = str(tmp_path / "ck.pt")
    torch.manual_seed(99)
    NN.save_checkpoint(path, gen, opt, 41)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert REF_KEYS <= set(ck.keys()) and set(ck.keys()) - REF_KEYS == {"msgm_sde"}
    assert set(ck["msgm_sde"]) == {"G", "L_G", "r_T"}
    gen2, opt2 = _objects(1000)  # different random G / r_T / weights
    assert not torch.equal(gen2.base_sde.G, gen.base_sde.G)
    assert NN.load_checkpoint(path, gen2, opt2, "cpu") == 41
    for k, v in gen.state_dict().items():
        assert torch.equal(gen2.state_dict()[k], v), k
    for n in ("G", "L_G", "r_T"):
        assert torch.equal(getattr(gen2.base_sde, n), getattr(gen.base_sde, n)), n
    assert torch.equal(torch.get_rng_state(), ck["torch_rng"])
    s1, s2 = opt.state_dict(), opt2.state_dict()
    for k in s1["state"]:
        for f in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(s1["state"][k][f], s2["state"][k][f])


def test_sparse_sde_tensors_and_reference_style_reader(tmp_path):
    torch.manual_seed(3)
    T = torch.nn.Parameter(torch.FloatTensor([1.0]), requires_grad=False)
    base = P.MSGMsde(torch.randn(64, 6), T=T, denseTensor=False, norm_map="log", device="cpu",
                     estim_cst_norm_dens_r_T=False)
    gen = P.PluginReverseSDE(base, P.MLP(6, premodule="NormalizeLogRadius"), T)
    opt = torch.optim.Adam(gen.parameters(), lr=1e-3)
    path = str(tmp_path / "ck.pt")
    NN.save_checkpoint(path, gen, opt, 0)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck["msgm_sde"]) == {"L_G", "r_T", "G_I", "G_J", "G_K", "G_V"}
    # the reference's loader (NN.py:21-39) touches exactly these keys and nothing else
    gen.load_state_dict(ck["model"])
    opt.load_state_dict(ck["optimizer"])
    assert ck["torch_rng"].dtype == torch.uint8 and ck["iteration"] == 0
    # persist_sde=False writes the reference's exact key set
    NN.save_checkpoint(path, gen, opt, 0, persist_sde=False)
    assert set(torch.load(path, map_location="cpu", weights_only=False).keys()) == REF_KEYS
