"""optim.Adam against torch.optim.Adam where the two could silently part: parameters without a gradient,
per-parameter step counters, ``param_groups`` lr edits, the table's row-form gradient, and ``main``'s FM warm start.
The reference builds ``torch.optim.Adam(params=model.parameters(), lr, weight_decay)`` (src/main/pretrain_main.py:181)."""
import numpy as np
import pytest
import torch

from oracle import torch_port as TP

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F, D, N = 15, 10, 300


def close(a, b, rtol=2e-5):
    a, b = a.detach().cpu().double().numpy(), b.detach().cpu().double().numpy()
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(float(np.abs(b).max()), 1e-30))


def _pair(name="FM", seed=3):
    """(our model on the GPU, the CPU port of the reference model) with equal parameters."""
    from rl_ctr_prediction_b200 import pretrain_main as PM
    torch.manual_seed(seed)
    port = TP.PortCTR(name, N, F, D).eval()
    with torch.no_grad():
        for k, p in port.named_parameters():
            if "embedding" in k or k == "linear.weight":
                p.mul_(0.1)
    m = PM.get_model(name, N, F, D)
    m.load_state_dict(port.state_dict())
    return m.to(DEV).eval(), port


def _batches(steps, B=64, seed=0):
    rng = np.random.default_rng(seed)
    return [(torch.as_tensor(rng.integers(0, N, size=(B, F))), torch.as_tensor((rng.random(B) < 0.3).astype(np.int64)).unsqueeze(1))
            for _ in range(steps)]


def _our_step(m, opt, x, y):
    p = m(x.to(DEV))
    tl = torch.nn.BCELoss()(p, y.to(DEV).float())
    m.zero_grad()
    tl.backward()
    opt.step()
    return tl


def test_lr_edited_through_param_groups_matches_torch():
    """The reference raises lr by 1e-4 per epoch with a NEW optimizer; an LR scheduler edits param_groups[0]['lr'] of a live
    one.  Both the touched rows and the lazily replayed ones must see each step's own lr."""
    from rl_ctr_prediction_b200 import optim
    m, port = _pair("FM")
    opt = optim.Adam(params=m.parameters(), lr=1e-3, weight_decay=1e-5)
    popt = torch.optim.Adam(port.parameters(), lr=1e-3, weight_decay=1e-5)
    lossf = torch.nn.BCELoss()
    for s, (x, y) in enumerate(_batches(6)):
        if s in (2, 4):
            for o in (opt, popt):
                o.param_groups[0]["lr"] *= 1.7
        ref = TP.ctr_train_step(port, popt, lossf, x, y)
        tl = _our_step(m, opt, x, y)
        assert abs(tl.item() - ref) <= 1e-5 * max(1.0, abs(ref))
    ref_sd = port.state_dict()
    for k, v in m.state_dict().items():                                 # every row: untouched ones were replayed lazily
        close(v, ref_sd[k])


def test_parameter_without_gradient_is_skipped_and_keeps_its_own_step():
    """torch.optim.Adam skips a parameter whose .grad is None and keeps state['step'] per parameter: one that gets its first
    gradient at the optimizer's third step takes ITS step 1 (bias correction 1 - beta^1)."""
    from rl_ctr_prediction_b200 import optim
    torch.manual_seed(0)
    a0, b0 = torch.randn(40, 7), torch.randn(33)
    ours = [torch.nn.Parameter(a0.clone().to(DEV)), torch.nn.Parameter(b0.clone().to(DEV))]
    ref = [torch.nn.Parameter(a0.clone()), torch.nn.Parameter(b0.clone())]
    opt = optim.Adam(ours, lr=1e-2, weight_decay=1e-3)
    ropt = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-3)
    for s in range(5):
        ga, gb = torch.randn(40, 7), torch.randn(33)
        for ps in (ours, ref):
            ps[0].grad = ga.clone().to(ps[0].device)
            ps[1].grad = gb.clone().to(ps[1].device) if s >= 2 and s != 3 else None      # late first gradient, then a gap
        opt.step()
        ropt.step()
        close(ours[0], ref[0], 1e-6)
        close(ours[1], ref[1], 1e-6)
    assert opt._dense_count == [5, 2] and opt._dense_steps.tolist() == [5, 2]


def test_table_without_backward_is_not_stepped():
    """No backward since the last step -> the table takes no step (no L2 decay either), like a parameter with .grad None."""
    from rl_ctr_prediction_b200 import optim
    m, port = _pair("FM")
    opt = optim.Adam(params=m.parameters(), lr=1e-3, weight_decay=1e-5)
    popt = torch.optim.Adam(port.parameters(), lr=1e-3, weight_decay=1e-5)
    lossf = torch.nn.BCELoss()
    bs = _batches(3)
    TP.ctr_train_step(port, popt, lossf, *bs[0])
    _our_step(m, opt, *bs[0])
    m.zero_grad()
    m.bias.grad = torch.zeros_like(m.bias)           # only the bias has a gradient in this step
    opt.step()
    popt.zero_grad(set_to_none=True)
    port.bias.grad = torch.zeros_like(port.bias)
    popt.step()
    assert m._opt.host_step == 1 and int(m._opt.step.item()) == 1
    TP.ctr_train_step(port, popt, lossf, *bs[2])
    _our_step(m, opt, *bs[2])
    ref_sd = port.state_dict()
    for k, v in m.state_dict().items():
        close(v, ref_sd[k])


def test_backward_without_rlctr_optimizer_raises():
    from rl_ctr_prediction_b200 import _lib
    m, _ = _pair("FM")
    x, y = _batches(1)[0]
    stock = torch.optim.Adam(m.parameters(), lr=1e-3)          # would silently skip the table: refuse instead
    p = m(x.to(DEV))
    tl = torch.nn.BCELoss()(p, y.to(DEV).float())
    with pytest.raises(_lib.RlctrError, match="optim.Adam"):
        tl.backward()
    del stock


def test_main_warm_starts_fnn_from_the_fm_checkpoint(tmp_path):
    """src/main/pretrain_main.py:164-166: FNN / IPNN / OPNN load '<campaign>FMbest.pth' into feature_embedding before training."""
    from rl_ctr_prediction_b200 import pretrain_main as PM
    rng = np.random.default_rng(0)
    root = str(tmp_path) + "/"
    import os
    os.makedirs(root + "ipinyou/1458/")
    n = 600
    ids = np.stack([rng.integers(f * 20, f * 20 + 20, size=n) for f in range(F)], axis=1)
    rows = np.column_stack([(rng.random(n) < 0.3).astype(np.int64), ids])
    np.savetxt(root + "ipinyou/1458/train.txt", rows, fmt="%d", delimiter=",")
    np.savetxt(root + "ipinyou/1458/day_index.csv", np.array([[6, 0, 199], [11, 200, 399], [12, 400, 599]]), fmt="%d", delimiter=",")
    args = dict(data_path=root, dataset_name="ipinyou/", campaign_id="1458/", valid_day=11, test_day=12, latent_dims=4,
                learning_rate=1e-3, weight_decay=1e-5, early_stop_type="loss", batch_size=128, device=DEV,
                save_param_dir=root + "params/")
    os.makedirs(root + "params/")
    with pytest.raises(FileNotFoundError, match="FMbest"):
        PM.main(model_name="FNN", epoch=0, **args)
    PM.setup_seed(1)
    PM.main(model_name="FM", epoch=1, **args)
    fm = torch.load(root + "params/1458/FMbest.pth")
    PM.main(model_name="FNN", epoch=0, **args)                       # epoch = 0: the checkpoint written is the warm start itself
    fnn = torch.load(root + "params/1458/FNNbest.pth")
    assert torch.equal(fnn["feature_embedding.weight"].cpu(), fm["feature_embedding.weight"].cpu())
