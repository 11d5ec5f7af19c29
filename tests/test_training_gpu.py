"""GPU parity of the training step (BASELINE.json configs[4]): loss and every parameter gradient of
`tqdne_b200.training.TrainStep1D` against autograd through the fp32 oracle of LightningEDM.step (tqdne/edm.py:115-134),
the Adam + EMA update against torch.optim.Adam / torch lerp, and the dropout / multi-step plumbing."""
import numpy as np
import pytest
import torch

from oracle import torch_ref
from oracle.weights import seeded_state_dict, shapes_of
from tests.conftest import rel_l2
from tests.helpers import unet_cfg

pytestmark = pytest.mark.gpu


def _edm(seed=31):
    import tqdne_b200 as tq

    edm = tq.LightningEDM(unet_cfg("1d"), {}, num_sampling_steps=18)
    sd = seeded_state_dict(shapes_of(edm), seed)
    edm.load_state_dict(sd)
    return edm.cuda(), sd


def _reference_loss_and_grads(sd, x, cond, sigma, noise):
    """LightningEDM.step (edm.py:115-134) on the fp32 oracle under autograd, explicit sigma / noise."""
    P = {k: v.detach().cuda().clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("time_embed.W"))
         for k, v in sd.items()}
    xn = x + noise * sigma[:, None, None]
    pred = torch_ref.denoise(P, unet_cfg("1d"), xn, sigma, cond)
    w = (sigma**2 + 0.25) / (sigma * 0.5) ** 2
    loss = ((pred - x) ** 2 * w[:, None, None]).mean()
    loss.backward()
    return loss.detach(), {k[len("unet."):]: v.grad for k, v in P.items() if v.requires_grad and v.grad is not None}


@pytest.mark.parametrize("N,L", [(2, 512), (1, 4064)], ids=["2x512", "1x4064"])
def test_training_step_loss_and_gradients_match_autograd(N, L):
    """(1, 4064) is the reference signal length: 508 attention tokens (ragged against the 128-row tiles of both attention
    backward kernels), 4064 / 2032 / 1016 / 508 positions per level."""
    from tqdne_b200.training import TrainStep1D

    edm, sd = _edm()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, 6, L, device="cuda", generator=g)
    cond = torch.randn(N, 5, device="cuda", generator=g)
    sigma = torch.tensor([0.4, 2.5], device="cuda")[:N]
    noise = torch.randn(N, 6, L, device="cuda", generator=g)
    step = TrainStep1D(edm, N, L, dropout=0.0)
    loss = step.forward_backward(x, cond, sigma=sigma, noise=noise)
    torch.cuda.synchronize()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref_loss, ref = _reference_loss_and_grads(sd, x, cond, sigma, noise)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert abs(float(loss) - float(ref_loss)) < 1e-2 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    got = step.grads_by_name()
    assert set(got) == set(ref), set(got) ^ set(ref)
    errs = {k: rel_l2(got[k], ref[k]) for k in ref}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    flat_g = torch.cat([got[k].flatten() for k in sorted(ref)])
    flat_r = torch.cat([ref[k].flatten() for k in sorted(ref)])
    total = rel_l2(flat_g, flat_r)
    cos = float(torch.nn.functional.cosine_similarity(flat_g, flat_r, dim=0))
    print(f"loss {float(loss):.6f} vs {float(ref_loss):.6f}; gradient rel-L2 {total:.3e}, cosine {cos:.6f}; worst {worst}")
    # bf16 activations AND bf16 activation gradients through ~80 layers: the whole gradient within 3e-2, no single
    # parameter tensor worse than 1.5e-1
    assert total < 3e-2 and cos > 0.999, (total, cos, worst)
    assert worst[0][1] < 1.5e-1, worst


def test_adam_ema_update_matches_torch_and_loss_decreases():
    from tqdne_b200.training import TrainStep1D

    N, L = 2, 512
    edm, _ = _edm(seed=32)
    step = TrainStep1D(edm, N, L, lr=1e-3, max_steps=50, dropout=0.1)
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(N, 6, L, device="cuda", generator=g)
    cond = torch.randn(N, 5, device="cuda", generator=g)
    sigma = torch.tensor([0.7, 1.5], device="cuda")
    noise = torch.randn(N, 6, L, device="cuda", generator=g)
    # one update against torch.optim.Adam on the same gradient
    p0 = step.store.P.clone()
    l0 = float(step.forward_backward(x, cond, sigma=sigma, noise=noise))
    grad = step.store.G.clone()
    tp = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([tp], lr=step.lr0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=0.0)
    tp.grad = grad.clone()
    opt.step()
    sched.step()
    step.optimizer_step()
    torch.cuda.synchronize()
    assert rel_l2(step.store.P - p0, tp.detach() - p0) < 1e-4
    assert rel_l2(step.store.EMA - p0, (tp.detach() - p0) * 0.001) < 5e-3   # 1e-3 * update on top of O(0.1) values in fp32
    # second update: the scheduler has stepped once
    p1 = step.store.P.clone()
    step.forward_backward(x, cond, sigma=sigma, noise=noise)
    tp.grad = step.store.G.clone()
    opt.step()
    sched.step()
    step.optimizer_step()
    assert rel_l2(step.store.P - p1, tp.detach() - p1) < 1e-4
    # a few more steps on the same batch: the loss must go down (dropout active, fresh masks per step)
    losses = [l0]
    for _ in range(6):
        losses.append(float(step.forward_backward(x, cond, sigma=sigma, noise=noise)))
        step.optimizer_step()
    assert np.isfinite(losses).all() and losses[-1] < 0.9 * losses[0], losses
    # masters written back to the module: the sampling path sees the trained weights
    step.sync_module()
    w = edm.unet.out[2].weight.detach()
    assert rel_l2(w, step.store.to_module_layout(step.store.P, edm.unet.out[2].weight)) < 1e-6


def test_lightning_module_training_step_entry_point():
    """LightningEDM.training_step(batch, batch_idx) (reference: edm.py:115-139) runs the engine's step for the 1D config and
    refuses the configurations it is not built for."""
    import tqdne_b200 as tq
    from tests.helpers import CFG, arch

    edm, _ = _edm(seed=33)
    edm.optimizer_params = {"learning_rate": 1e-3, "max_steps": 10, "eta_min": 0.0}
    g = torch.Generator(device="cuda").manual_seed(9)
    batch = {"signal": torch.randn(2, 6, 512, device="cuda", generator=g), "cond": torch.randn(2, 5, device="cuda", generator=g)}
    l0 = float(edm.training_step(batch, 0))
    w0 = edm.unet.out[2].weight.detach().clone()
    for i in range(3):
        edm.training_step(batch, i + 1)
    assert np.isfinite(l0)
    edm.sync_trained_weights()
    assert float((edm.unet.out[2].weight.detach() - w0).abs().max()) > 0       # the module sees the trained weights
    out = edm.sample((2, 6, 512), cond=batch["cond"])                        # and the sampling path still runs on them
    assert out.shape == (2, 6, 512) and bool(torch.isfinite(out).all())
    enc_cfg, dec_cfg = arch().get_2d_autoencoder_configs(CFG)
    latent = tq.LightningEDM(unet_cfg("latent2d"), {}, autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {})).cuda()
    with pytest.raises(NotImplementedError):
        latent.training_step({"signal": torch.zeros(1, 3, 128, 128, device="cuda")}, 0)


def test_training_step_matches_reference_golden():
    """The engine's step against the UNMODIFIED reference's LightningEDM.step under autograd (tests/golden/
    train_step_1d.npz, oracle/make_golden_train.py): same weights, batch, sigma and noise."""
    from tests.conftest import GOLDEN
    from tqdne_b200.training import TrainStep1D

    z = np.load(GOLDEN / "train_step_1d.npz")
    edm, _ = _edm(seed=int(z["seed"]))
    x, cond, sigma, noise = (torch.from_numpy(z[k]).cuda() for k in ("x", "cond", "sigma", "noise"))
    step = TrainStep1D(edm, x.shape[0], x.shape[2], dropout=0.0)
    loss = float(step.forward_backward(x, cond, sigma=sigma, noise=noise))
    assert abs(loss - float(z["loss"])) < 1e-2 * float(z["loss"]), (loss, float(z["loss"]))
    got = step.grads_by_name()
    names = [str(n) for n in z["grad_names"]]
    assert set(names) == {"unet." + k for k in got}
    norms = torch.tensor([float(got[n[len("unet."):]].norm()) for n in names], dtype=torch.float64)
    assert rel_l2(norms, torch.from_numpy(z["grad_norms"])) < 2e-2
    for k in z.files:
        if k.startswith("grad:"):
            assert rel_l2(got[k[len("grad:unet."):]].cpu(), torch.from_numpy(z[k])) < 5e-2, k


def test_training_state_survives_a_change_of_batch_shape():
    """A ragged last batch (or an evaluation batch) builds a new static tape: the parameters, Adam moments, EMA and the
    position on the cosine schedule must carry over, and a tape that sat idle while another one stepped the optimiser
    must refresh its bf16 operand copies before it runs again."""
    edm, _ = _edm(seed=33)
    g = torch.Generator(device="cuda").manual_seed(9)

    def batch(n, L):
        return {"signal": torch.randn(n, 6, L, device="cuda", generator=g), "cond": torch.randn(n, 5, device="cuda", generator=g)}

    a, b = batch(2, 512), batch(1, 512)
    edm.training_step(a)
    ta = edm._train_step(a)
    m_norm = float(ta.store.M.norm())
    p_after_1 = ta.store.P.clone()
    assert ta.step_count == 1 and m_norm > 0
    edm.training_step(b)                       # another batch size: new tape, same state
    tb = edm._train_step(b)
    assert tb is not ta and tb.store is ta.store and tb.step_count == 2
    assert float((tb.store.P - p_after_1).norm()) > 0 and float(tb.store.M.norm()) > 0
    assert ta.copies_version != ta.store.version          # tape A's operand copies are stale now ...
    edm.training_step(a)
    assert edm._train_step(a) is ta and ta.step_count == 3
    assert ta.copies_version == ta.store.version          # ... and were refreshed before / after it ran
    edm.sync_trained_weights()
    w = edm.unet.out[2].weight.detach().float()
    assert rel_l2(ta.store.to_module_layout(ta.store.P, edm.unet.out[2].weight), w) < 1e-7


def test_training_tape_graph_replay_equals_eager_passes():
    """From the third pass on the tape is ONE CUDA-graph replay (all launch arguments are static; dropout decisions come
    from a device counter): with dropout off and the same explicit sigma / noise it must give the loss and gradients of the
    eager passes (up to the order of the fp32 atomics in the weight-gradient split-K)."""
    from tqdne_b200.training import TrainStep1D

    edm, _ = _edm(seed=35)
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(2, 6, 512, device="cuda", generator=g)
    cond = torch.randn(2, 5, device="cuda", generator=g)
    sigma = torch.tensor([0.7, 3.0], device="cuda")
    noise = torch.randn(2, 6, 512, device="cuda", generator=g)
    step = TrainStep1D(edm, 2, 512, dropout=0.0)
    losses, grads = [], []
    for _ in range(4):
        losses.append(float(step.forward_backward(x, cond, sigma=sigma, noise=noise)))
        grads.append(step.store.G.clone())
    # two graphs: the tape is split where the late gradient bucket is final (the data-parallel step starts its all-reduce there)
    assert isinstance(step._graph, list) and len(step._graph) == 2 and all(isinstance(g_, torch.cuda.CUDAGraph) for g_ in step._graph), \
        "the third pass should have captured the tape"
    assert 0 < step.bwd_split < len(step.bwd) and 0 < step.store.tail_off < step.store.n
    assert abs(losses[3] - losses[1]) < 1e-6 * abs(losses[1]) and abs(losses[2] - losses[1]) < 1e-6 * abs(losses[1])
    assert rel_l2(grads[3], grads[1]) < 1e-5 and rel_l2(grads[2], grads[1]) < 1e-5
