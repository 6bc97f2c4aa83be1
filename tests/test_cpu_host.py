"""CPU-side (no GPU) tests: the C-ABI library loads and exports every declared symbol, the
host-side mirror of the reference API behaves like the reference (state_dict layout, from_pth,
grid size, registry, error behaviour)."""
import ctypes
import os
import re

import pytest
import torch

import yogo_b200
from yogo_b200 import _lib as L
from yogo_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "yogo_b200.h")).read()
    declared = set(re.findall(r"\b(yg_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"yg_status"}
    lib = ctypes.CDLL(L.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(L.EXPORTED_SYMBOLS) == declared
    assert L.load().yg_version() == 100


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = yogo_b200.YOGO((64, 96), 0.05, 0.05, 7)
    with pytest.raises(L.YogoB200Error):
        m(torch.zeros(1, 1, 64, 96))
    with pytest.raises(L.YogoB200Error):
        yogo_b200.YOGOLoss()(torch.zeros(1, 12, 4, 4), torch.zeros(1, 6, 4, 4))
    with pytest.raises(L.YogoB200Error):
        yogo_b200.format_preds(torch.zeros(12, 4, 4))


def test_format_preds_argument_errors():
    # /root/reference/yogo/utils/prediction_formatting.py:52-60
    with pytest.raises(ValueError):
        yogo_b200.format_preds(torch.zeros(1, 12, 4, 4))
    with pytest.raises(ValueError):
        yogo_b200.format_preds(torch.zeros(12, 4, 4), box_format="bogus")


def test_count_cells_known_answers():
    # /root/reference/tests/test_count_predictions.py
    f = yogo_b200.count_cells_for_formatted_preds
    inp = torch.zeros(3, 5)
    inp[:, 0] = 1
    assert f(inp).tolist() == [3, 0, 0, 0, 0]
    row = torch.tensor([0.1, 0.2, 0.3, 0.4])
    assert f(torch.stack([row, row, row])).tolist() == [0, 0, 0, 3]
    inp = torch.tensor([[0.2, 0.4, 0.2, 0.2]] * 3)
    assert f(inp, min_confidence_threshold=0.6).tolist() == [0, 0, 0, 0]
    inp = torch.tensor([[0.2, 0.7, 0.2, 0.2], [0.2, 0.4, 0.2, 0.2], [0.2, 0.4, 0.9, 0.2]])
    assert f(inp, min_confidence_threshold=0.6).tolist() == [0, 1, 1, 0]
    with pytest.raises(ValueError):
        f(torch.zeros(3))
    with pytest.raises(ValueError):
        f(torch.zeros(3, 4), min_confidence_threshold=2)


def test_grid_size_and_buffers():
    m = yogo_b200.YOGO((772, 1032), 0.0425, 0.0555, 7)
    assert (m.Sx, m.Sy) == (129, 97)  # docs/recipes.md:133
    assert m.num_params() == 541852
    sd = m.state_dict()
    for k, shape in [("img_size", (2,)), ("_Cxs", (97, 129)), ("_Cys", (97, 129)), ("model.0.0.weight", (16, 1, 3, 3)),
                     ("model.7.weight", (12, 128, 1, 1)), ("model.5.1.running_var", (128,))]:
        assert tuple(sd[k].shape) == shape, k
    assert sd["num_classes"].dtype == torch.int64 and sd["is_rgb"].dtype == torch.bool
    m.resize_model(img_height=193)
    assert (m.Sx, m.Sy) == (129, 25)
    assert float(m.height_multiplier) == pytest.approx(772 / 193)
    assert tuple(m._Cys.shape) == (25, 129)
    assert yogo_b200.YOGO((772, 1032), 0.05, 0.05, 7, model_func=yogo_b200.get_model_func("double_filters")).num_params() == 2161964


@pytest.mark.parametrize("name", ["base_model", "silu_model"])
def test_model_io_roundtrip(tmp_path, name):
    # /root/reference/tests/test_model_io.py:38-57
    y = yogo_b200.YOGO(img_size=(772, 1032), anchor_w=0.05, anchor_h=0.05, num_classes=7,
                       model_func=yogo_b200.get_model_func(name))
    path = tmp_path / "test.pth"
    torch.save({"epoch": 0, "step": 0, "model_state_dict": y.state_dict(), "model_version": y.model_version}, str(path))
    z, meta = yogo_b200.YOGO.from_pth(path)
    assert z.model_version == name and meta["step"] == 0
    for k in ("anchor_w", "anchor_h", "num_classes", "is_rgb", "normalize_images", "clip_value",
              "height_multiplier", "width_multiplier"):
        assert getattr(y, k) == getattr(z, k)
    for p1, p2 in zip(y.parameters(), z.parameters()):
        assert p1.data.ne(p2.data).sum() == 0
    # legacy checkpoints lacking newer buffers still load (model.py:112-126)
    sd = y.state_dict()
    for k in ("is_rgb", "clip_value", "height_multiplier", "width_multiplier", "normalize_images"):
        sd.pop(k)
    torch.save({"model_state_dict": sd, "normalize_images": True}, str(path))
    z, meta = yogo_b200.YOGO.from_pth(path, inference=True)
    assert z.model_version == "base_model" and bool(meta["normalize_images"]) and not z.training


def test_registry_and_plan_compiler():
    assert yogo_b200.get_model_func(None) is yogo_b200.MODELS["base_model"]
    assert yogo_b200.get_model_func("nope") is yogo_b200.MODELS["base_model"]

    @yogo_b200.register_model
    def my_tiny(num_classes, rgb_input=False):
        from torch import nn
        return nn.Sequential(nn.Sequential(nn.Conv2d(1, 8, 3, padding=1), nn.SiLU()), nn.Conv2d(8, 5 + num_classes, 1))

    assert "my_tiny" in yogo_b200.MODELS
    plan = engine.compile_plan(my_tiny(3))
    assert len(plan.blocks) == 1 and plan.blocks[0].act == L.ACT_SILU
    del yogo_b200.MODELS["my_tiny"]
    from torch import nn
    with pytest.raises(NotImplementedError):
        engine.compile_plan(nn.Sequential(nn.Sequential(nn.Conv2d(1, 8, 5, padding=2)), nn.Conv2d(8, 12, 1)))
    with pytest.raises(NotImplementedError):
        engine.compile_plan(nn.Sequential(nn.Sequential(nn.Conv2d(1, 8, 3, padding=1), nn.ReLU()), nn.Conv2d(8, 12, 1)))


def test_trainer_optimizer_state_round_trip(tmp_path):
    """DataParallelTrainer.state_dict / checkpoint: the reference's checkpoint dict (train.py:266-293) with an
    optimizer_state_dict that torch.optim.AdamW itself accepts; a resumed trainer continues the moments and the schedule."""
    import yogo_b200
    from yogo_b200.train import DataParallelTrainer

    torch.manual_seed(0)
    net = yogo_b200.YOGO((64, 96), 0.05, 0.05, 7)
    tr = DataParallelTrainer(net, total_steps=100)
    tr.exp_avg.normal_()
    tr.exp_avg_sq.uniform_(0, 1)
    tr.step_count = 7
    path = tmp_path / "ckpt.pth"
    tr.checkpoint(path, "base_model", epoch=3, classes=["a"] * 7)
    ck = torch.load(path, weights_only=False)
    assert {"epoch", "step", "normalize_images", "classes", "model_name", "model_state_dict", "optimizer_state_dict",
            "model_version"} <= set(ck)
    assert ck["step"] == 7 and ck["epoch"] == 3
    # torch's own optimizer loads it and sees the same moments, parameter by parameter
    opt = torch.optim.AdamW(net.parameters(), lr=3e-4, weight_decay=5e-2)
    opt.load_state_dict(ck["optimizer_state_dict"])
    for p in net.parameters():
        o, k = tr.offsets[id(p)]
        assert torch.equal(opt.state[p]["exp_avg"].reshape(-1), tr.exp_avg[o:o + k])
        assert float(opt.state[p]["step"]) == 7.0
    assert abs(opt.param_groups[0]["lr"] - tr.lr_at(7)) < 1e-12
    # and the other direction: a state dict produced by torch.optim.AdamW resumes a fresh trainer
    net2, _ = yogo_b200.YOGO.from_pth(path)
    tr2 = DataParallelTrainer(net2, total_steps=100)
    tr2.load_optimizer_state_dict(opt.state_dict())
    assert tr2.step_count == 7
    for p, p2 in zip(net.parameters(), net2.parameters()):
        (o, k), (o2, k2) = tr.offsets[id(p)], tr2.offsets[id(p2)]
        assert torch.equal(tr.exp_avg_sq[o:o + k], tr2.exp_avg_sq[o2:o2 + k2])
    # parameters that stop aliasing the flat buffer are detected
    net2.model[0][0].weight.data = net2.model[0][0].weight.data.clone()
    with pytest.raises(RuntimeError, match="no longer aliases"):
        tr2._check_views()


def test_collate_batch_robust_cpu():
    """yogo/data/utils.py:49-63 without a device: pure stacking, identity transforms."""
    from yogo_b200.data import collate_batch_robust
    a = (torch.zeros(1, 4, 4, dtype=torch.uint8), torch.zeros(6, 2, 2))
    b = (torch.ones(1, 4, 4, dtype=torch.uint8), torch.ones(6, 2, 2))
    x, y = collate_batch_robust([a, None, b])
    assert tuple(x.shape) == (2, 1, 4, 4) and tuple(y.shape) == (2, 6, 2, 2) and int(x[1].sum()) == 16
    with pytest.raises(ValueError):
        collate_batch_robust([None])


def test_gradient_clamp_keeps_the_constructor_clip_value(tmp_path):
    """model.py:76-77: the reference's hook closes over the constructor argument; loading a checkpoint whose `clip_value`
    buffer differs (from_pth builds with the default 1.0) must not change the clamp the kernels apply."""
    import yogo_b200
    net = yogo_b200.YOGO((64, 96), 0.05, 0.05, 7, clip_value=0.25)
    assert net._clip_value_f == 0.25
    path = tmp_path / "c.pth"
    torch.save({"model_state_dict": net.state_dict(), "model_version": "base_model"}, path)
    loaded, _ = yogo_b200.YOGO.from_pth(path)
    assert float(loaded.clip_value) == 0.25 and loaded._clip_value_f == 1.0
