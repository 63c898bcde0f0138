"""Policy inference in the rollout loop: zb_actor_f32 against the CPU restatement of the reference's
Actor.forward (/root/reference/train.py:946-968), and the restatement against an independent implementation of the same published
layers (torch.nn.Linear / torch.nn.GRUCell on the CPU)."""
import ctypes

import numpy as np
import pytest

import ksim_gym_zbot_b200  # noqa: F401
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.policy import HIDDEN, ActorParams, ZbActor


def _inputs(n, depth, seed):
    g = np.random.default_rng(seed)
    obs = g.normal(0, 1.0, (n, 50)).astype(np.float32)
    carry = g.uniform(-1, 1, (n, depth, HIDDEN)).astype(np.float32)
    return obs, carry, g.uniform(0, 1, (n, M.NU)).astype(np.float32), g.normal(0, 1, (n, M.NU)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------------
# CPU: the oracle
# ---------------------------------------------------------------------------------------------------------------------------
def test_oracle_matches_torch_layers():
    """eqx.nn.GRUCell == torch.nn.GRUCell with bias_ih = bias and bias_hh = [0, 0, bias_n] (same gate order and update rule)."""
    import torch
    import zb_policy_oracle as PO

    prm = ActorParams.random(seed=3)
    a = prm.arrays
    obs, carry, _, _ = _inputs(64, prm.depth, 5)
    with torch.no_grad():
        lin_in, lin_out = torch.nn.Linear(50, HIDDEN).double(), torch.nn.Linear(HIDDEN, 300).double()
        lin_in.weight.copy_(torch.tensor(a["input_proj.weight"])); lin_in.bias.copy_(torch.tensor(a["input_proj.bias"]))
        lin_out.weight.copy_(torch.tensor(a["output_proj.weight"])); lin_out.bias.copy_(torch.tensor(a["output_proj.bias"]))
        x = lin_in(torch.tensor(obs).double())
        carries = []
        for i in range(prm.depth):
            cell = torch.nn.GRUCell(HIDDEN, HIDDEN).double()
            cell.weight_ih.copy_(torch.tensor(a[f"rnns.{i}.weight_ih"])); cell.weight_hh.copy_(torch.tensor(a[f"rnns.{i}.weight_hh"]))
            cell.bias_ih.copy_(torch.tensor(a[f"rnns.{i}.bias"]))
            cell.bias_hh.zero_(); cell.bias_hh[2 * HIDDEN:].copy_(torch.tensor(a[f"rnns.{i}.bias_n"]))
            x = cell(x, torch.tensor(carry[:, i]).double())
            carries.append(x)
        out = lin_out(x).numpy()
        std_t = torch.clamp((torch.nn.functional.softplus(torch.tensor(out[:, 100:200])) + 0.01) * 1.0, max=1.0).numpy()
    mean, std, logits, nc = PO.actor_forward(prm.oracle_dict(), obs, carry)
    np.testing.assert_allclose(nc, np.stack([c.numpy() for c in carries], 1), rtol=0, atol=1e-12)
    np.testing.assert_allclose(mean, out[:, :100].reshape(-1, 20, 5) + M.BIAS_POSE[None, :, None].astype(np.float32), rtol=0, atol=1e-12)
    np.testing.assert_allclose(std, std_t.reshape(-1, 20, 5), rtol=0, atol=1e-12)
    np.testing.assert_allclose(logits, out[:, 200:].reshape(-1, 20, 5), rtol=0, atol=1e-12)


def test_oracle_action_selection():
    import zb_policy_oracle as PO

    g = np.random.default_rng(0)
    mean, std, logits = g.normal(size=(7, 20, 5)), g.uniform(0.1, 1, (7, 20, 5)), g.normal(size=(7, 20, 5))
    a = PO.select_action(mean, std, logits)
    k = logits.argmax(-1)
    assert np.array_equal(a, np.take_along_axis(mean, k[..., None], -1)[..., 0])  # mode: mean of the most probable component
    # sample: u sweeps the CDF -> every component is reachable, and eps = 0 returns that component's mean
    p = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    for j in range(5):
        u = p[..., :j].sum(-1) + 0.5 * p[..., j]
        s = PO.select_action(mean, std, logits, u, np.zeros((7, 20)))
        np.testing.assert_allclose(s, mean[..., j], atol=1e-12)
    s = PO.select_action(mean, std, logits, np.zeros((7, 20)), np.ones((7, 20)))
    np.testing.assert_allclose(s, mean[..., 0] + std[..., 0], atol=1e-12)
    assert PO.flops_per_env() == 2 * (50 * 128 + 5 * 6 * 128 * 128 + 128 * 300)


def test_pack_layout_and_abi():
    prm = ActorParams.random(seed=1, depth=3)
    pk = prm.pack()
    w = prm.arrays["rnns.2.weight_hh"]
    for (k, g, f) in ((0, 0, 0), (5, 1, 7), (127, 2, 127), (64, 2, 3)):
        assert pk["gru_whh_t"][2, k, g, f] == w[g * HIDDEN + f, k]
    assert pk["in_wt"].shape == (50, HIDDEN) and pk["in_wt"][3, 9] == prm.arrays["input_proj.weight"][9, 3]
    assert pk["out_wt"].shape == (HIDDEN, 300) and pk["gru_b"].shape == (3, 3, HIDDEN)
    # tensor-core operand tiles: [layer][k-tile t][matrix m][k-chunk c][f][j] = W_m[f][8 t + 4 c + j], m = (W_ih r, z, n, W_hh r, z, n)
    assert pk["gru_w_tc"].shape == (3, 16, 6, 2, HIDDEN, 4)
    wi = prm.arrays["rnns.1.weight_ih"]
    for (t, m, c, f, j) in ((0, 0, 0, 0, 0), (3, 2, 1, 17, 2), (15, 4, 1, 127, 3), (7, 5, 0, 64, 1)):
        src = (wi if m < 3 else prm.arrays["rnns.1.weight_hh"])[(m % 3) * HIDDEN + f, 8 * t + 4 * c + j]
        assert pk["gru_w_tc"][1, t, m, c, f, j] == src
    assert pk["out_w_tc"].shape == (16, 3, 2, HIDDEN, 4)
    wo = prm.arrays["output_proj.weight"]
    assert pk["out_w_tc"][5, 2, 1, 43, 2] == wo[2 * HIDDEN + 43, 8 * 5 + 4 * 1 + 2] and pk["out_w_tc"][0, 2, 0, 44, 0] == 0.0  # row 300: padding
    lib = ctypes.CDLL(ksim_gym_zbot_b200.engine.LIB_PATH)
    assert lib.zb_sizeof_actor() == ctypes.sizeof(ZbActor)
    assert hasattr(lib, "zb_actor_f32")
    with pytest.raises(ValueError):
        ActorParams({k: v for k, v in prm.arrays.items() if k != "rnns.1.bias_n"}, depth=3)


# ---------------------------------------------------------------------------------------------------------------------------
# GPU: kernel vs oracle
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1], ids=["fp32", "tf32x3-tcgen05"])
@pytest.mark.parametrize("n,depth,mix", [(1, 5, 5), (33, 5, 5), (1000, 5, 5), (4096, 5, 5), (77, 2, 3), (40, 1, 1)])
def test_actor_kernel_matches_oracle(n, depth, mix, precision):
    """Both GRU paths — fp32 on the CUDA cores, 3xTF32 on the tcgen05 tensor cores — against the fp64 restatement, with the fp32
    restatement's own deviation as the yard-stick (the 3xTF32 products keep ~21 mantissa bits per factor: same bound)."""
    import torch
    import zb_policy_oracle as PO
    from ksim_gym_zbot_b200.policy import ActorPolicy

    prm = ActorParams.random(seed=11 + depth, depth=depth, num_mixtures=mix)
    obs, carry, u, eps = _inputs(n, depth, 100 + n)
    pol = ActorPolicy(prm, n, "cuda:0", precision=precision)
    pol.carry.copy_(torch.tensor(carry))
    dist = torch.zeros((n, 3, M.NU, mix), device="cuda:0")
    act = torch.zeros((n, M.NU), device="cuda:0")
    d_obs = torch.tensor(obs, device="cuda:0")
    pol.enqueue(d_obs, act, dist_out=dist)  # mode
    torch.cuda.synchronize()
    mean, std, logits, nc = PO.actor_forward(prm.oracle_dict(), obs, carry)
    # fp32 kernel against the fp64 restatement; the fp32 restatement (numpy, other summation order) is the yard-stick
    m32, s32, l32, nc32 = PO.actor_forward(prm.oracle_dict(), obs, carry, dtype=np.float32)
    tol = lambda a64, a32: 1e-5 + 4 * float(np.abs(a32 - a64).max())  # noqa: E731
    np.testing.assert_allclose(pol.carry.cpu().numpy(), nc, rtol=0, atol=tol(nc, nc32))
    d = dist.cpu().numpy()
    np.testing.assert_allclose(d[:, 0], mean, rtol=0, atol=tol(mean, m32))
    np.testing.assert_allclose(d[:, 1], std, rtol=0, atol=tol(std, s32))
    np.testing.assert_allclose(d[:, 2], logits, rtol=0, atol=tol(logits, l32))
    # action selection on the KERNEL's own distribution parameters is exact (ties in the argmax aside, none with random weights)
    np.testing.assert_array_equal(act.cpu().numpy(), PO.select_action(d[:, 0], d[:, 1], d[:, 2]))
    # sample path from the same carry
    pol.carry.copy_(torch.tensor(carry))
    pol.enqueue(d_obs, act, noise_u=torch.tensor(u, device="cuda:0"), noise_eps=torch.tensor(eps, device="cuda:0"), dist_out=dist)
    torch.cuda.synchronize()
    d = dist.cpu().numpy()
    want = PO.select_action(d[:, 0].astype(np.float64), d[:, 1].astype(np.float64), d[:, 2].astype(np.float64), u, eps)
    got = act.cpu().numpy()
    close = np.abs(got - want) < 1e-5
    assert close.mean() > 0.999, close.mean()  # (a component boundary of the CDF within rounding of u may flip an element)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1], ids=["fp32", "tf32x3-tcgen05"])
def test_actor_recurrence_reset_and_errors(precision):
    import torch
    import zb_policy_oracle as PO
    from ksim_gym_zbot_b200.engine import ZB_EINVAL, ZB_EUNSUP, ZbError
    from ksim_gym_zbot_b200.policy import ActorPolicy

    n, T = 96, 12
    prm = ActorParams.random(seed=2)
    pol = ActorPolicy(prm, n, "cuda:0", precision=precision)
    g = np.random.default_rng(4)
    obs = g.normal(0, 1, (T, n, 50)).astype(np.float32)
    done = (g.uniform(size=(T, n)) < 0.1).astype(np.uint8)
    carry = np.zeros((n, prm.depth, HIDDEN))
    for t in range(T):
        a = pol(torch.tensor(obs[t], device="cuda:0"), carry_reset=None if t == 0 else torch.tensor(done[t - 1], device="cuda:0"))
        if t > 0:
            carry[done[t - 1] != 0] = 0.0  # the episode ended in step t-1: initial carry (train.py:1731-1735)
        mean, std, logits, carry = PO.actor_forward(prm.oracle_dict(), obs[t], carry)
        np.testing.assert_allclose(pol.carry.cpu().numpy(), carry, rtol=0, atol=2e-5)
        np.testing.assert_allclose(a.cpu().numpy(), PO.select_action(mean, std, logits), rtol=0, atol=1e-4)
    # in-place carry == separate buffers; a CUDA-graph replay of a step gives the same bits
    pol.carry.copy_(torch.tensor(carry.astype(np.float32)))
    o = torch.tensor(obs[0], device="cuda:0")
    ref = pol(o).clone()
    ref_carry = pol.carry.clone()
    pol.carry.copy_(torch.tensor(carry.astype(np.float32)))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            pol.enqueue(o, pol.action)
    torch.cuda.current_stream().wait_stream(s)
    pol.carry.copy_(torch.tensor(carry.astype(np.float32)))
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(pol.action, ref) and torch.equal(pol.carry, ref_carry)
    # error behaviour
    with pytest.raises(ZbError) as ei:
        pol.enqueue(o, pol.action, noise_u=torch.zeros((n, 20), device="cuda:0"))
    assert ei.value.code == ZB_EINVAL
    pol.desc.num_mixtures = 9
    with pytest.raises(ZbError) as ei:
        pol.enqueue(o, pol.action)
    assert ei.value.code == ZB_EUNSUP


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1], ids=["fp32", "tf32x3-tcgen05"])
def test_actor_kernel_is_deterministic_and_batch_independent(precision):
    """Races in the asynchronous pipelines (TMA ring, mbarrier phases, tcgen05 commit / TMEM reuse) would show up as run-to-run differences:
    40 launches from the same inputs are bitwise identical, and an environment's result does not depend on the batch it sits in (ragged
    tail, another position in the CTA grid)."""
    import torch
    from ksim_gym_zbot_b200.policy import ActorPolicy

    n = 4096 + 7
    prm = ActorParams.random(seed=5)
    obs, carry, u, eps = _inputs(n, prm.depth, 9)
    d_obs, d_carry = torch.tensor(obs, device="cuda:0"), torch.tensor(carry, device="cuda:0")
    d_u, d_eps = torch.tensor(u, device="cuda:0"), torch.tensor(eps, device="cuda:0")
    pol = ActorPolicy(prm, n, "cuda:0", precision=precision)
    ref_a = ref_c = None
    for rep in range(40):
        pol.carry.copy_(d_carry)
        pol.enqueue(d_obs, pol.action, noise_u=d_u, noise_eps=d_eps)
        if rep == 0:
            ref_a, ref_c = pol.action.clone(), pol.carry.clone()
        else:
            assert torch.equal(pol.action, ref_a) and torch.equal(pol.carry, ref_c), f"launch {rep} differs from launch 0"
    torch.cuda.synchronize()
    # the last 39 environments alone (other CTA index, other position inside the tile): same bits
    m = 39
    pol2 = ActorPolicy(prm, m, "cuda:0", precision=precision)
    pol2.carry.copy_(d_carry[n - m:])
    pol2.enqueue(d_obs[n - m:].contiguous(), pol2.action, noise_u=d_u[n - m:].contiguous(), noise_eps=d_eps[n - m:].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(pol2.action, ref_a[n - m:]) and torch.equal(pol2.carry, ref_c[n - m:])


def _golden():
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "actor_golden.npz"))
    return g, ActorParams.random(seed=int(g["seed"]))


def test_oracle_matches_actor_golden():
    """tests/golden/actor_golden.npz (written by tests/golden/make_actor_golden.py from torch.nn layers): the restatement reproduces three
    recurrent steps incl. the carry restarts."""
    import zb_policy_oracle as PO

    g, prm = _golden()
    carry = g["carry0"].astype(np.float64)
    for t in range(g["obs"].shape[0]):
        carry[g["reset"][t] != 0] = 0.0
        mean, std, logits, carry = PO.actor_forward(prm.oracle_dict(), g["obs"][t], carry)
        for got, key in ((mean, "means"), (std, "stds"), (logits, "logits"), (carry, "carries")):
            np.testing.assert_allclose(got, g[key][t], rtol=0, atol=2e-7, err_msg=f"{key} step {t}")  # (the fixture is stored in float32)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1], ids=["fp32", "tf32x3-tcgen05"])
def test_actor_kernel_matches_golden(precision):
    import torch
    from ksim_gym_zbot_b200.policy import ActorPolicy

    g, prm = _golden()
    T, n = g["obs"].shape[:2]
    pol = ActorPolicy(prm, n, "cuda:0", precision=precision)
    pol.carry.copy_(torch.tensor(g["carry0"]))
    dist = torch.zeros((n, 3, M.NU, prm.num_mixtures), device="cuda:0")
    for t in range(T):
        pol.enqueue(torch.tensor(g["obs"][t], device="cuda:0"), pol.action, carry_reset=torch.tensor(g["reset"][t], device="cuda:0"), dist_out=dist)
        torch.cuda.synchronize()
        d = dist.cpu().numpy()
        tol = 3e-6 if precision == 0 else 1e-5
        np.testing.assert_allclose(pol.carry.cpu().numpy(), g["carries"][t], rtol=0, atol=tol, err_msg=f"carry step {t}")
        np.testing.assert_allclose(d[:, 0], g["means"][t], rtol=0, atol=tol, err_msg=f"means step {t}")
        np.testing.assert_allclose(d[:, 1], g["stds"][t], rtol=0, atol=tol, err_msg=f"stds step {t}")
        np.testing.assert_allclose(d[:, 2], g["logits"][t], rtol=0, atol=tol, err_msg=f"logits step {t}")
