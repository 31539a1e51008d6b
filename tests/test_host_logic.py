"""Host-side logic that needs no GPU: resize-size rule, weight packing, generator tree / init parity with the
reference (golden), schedulers, loud failure without CUDA."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_resized_hw_matches_oracle_rule():
    from oracle import splice_ref as R
    from splice_b200 import _lib

    for (h, w) in [(224, 224), (213, 213), (128, 128), (448, 448), (900, 1200), (1200, 900), (982, 1280), (225, 300),
                   (100, 1000), (1000, 100), (223, 225), (481, 479)]:
        for size in (224, 448, 112):
            oh, ow = C.c_int(), C.c_int()
            _lib.splice_resized_hw(h, w, size, 480, C.byref(oh), C.byref(ow))
            assert (oh.value, ow.value) == R.resized_hw(h, w, size, 480), (h, w, size)


def test_pack_order_covers_every_dino_tensor():
    from oracle import dino_vit
    from splice_b200 import _lib
    from splice_b200.engine import DINO_ARCH, pack_vit_weights, packed_key_order

    for name in ("dino_vits16", "dino_vitb8"):
        sd = dino_vit.build(name).state_dict()
        keys = packed_key_order()
        assert sorted(keys) == sorted(sd.keys()) and len(keys) == 150
        packed = pack_vit_weights(sd, "cpu")
        patch, dim, heads = DINO_ARCH[name]
        d = _lib.SpliceVitDesc(patch, dim, heads, 12, sd["pos_embed"].shape[1], 1e-6)
        assert packed.numel() == _lib.splice_vit_packed_floats(C.byref(d))
        # spot-check the layout: cls first, final norm bias last
        assert torch.equal(packed[:dim], sd["cls_token"].reshape(-1))
        assert torch.equal(packed[-dim:], sd["norm.bias"])


def test_random_dino_state_dict_has_dino_shapes():
    from oracle import dino_vit
    from splice_b200.dino_init import random_dino_state_dict

    ref = dino_vit.build("dino_vits8").state_dict()
    sd = random_dino_state_dict("dino_vits8")
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref.items()}


def test_generator_tree_and_init_match_reference(golden_dir):
    """Same state_dict keys, parameter order and (seed 0) bit-identical initial weights as the reference's
    define_G('xavier', 0.02) — the golden was produced by the reference itself (oracle/make_golden.py)."""
    from oracle import splice_ref as R
    from splice_b200.models.networks import define_G

    gold = torch.load(golden_dir / "step_s16.pt")["netG"]
    torch.manual_seed(0)
    net = define_G("xavier", 0.02)
    sd = net.state_dict()
    assert list(sd.keys()) == list(gold.keys())
    for k in sd:
        assert torch.equal(sd[k], gold[k]), k
    assert [k for k, _ in net.named_parameters()] == R.generator_param_keys()
    assert sum(p.numel() for p in net.parameters()) == 1_037_523


def test_out_of_scope_generator_options_raise():
    from splice_b200.generator import NativeSkip
    from splice_b200.generator_x import NativeSkipX
    from splice_b200.models.unet.skip import skip

    with pytest.raises(NotImplementedError):
        skip(downsample_mode="lanczos2")
    with pytest.raises(NotImplementedError):
        skip(act_fun="Swish")
    with pytest.raises(NotImplementedError):
        skip(upsample_mode="nearest")
    with pytest.raises(NotImplementedError):
        skip(num_channels_skip=[4, 4, 0, 4, 4])
    # the default arguments build the engine tuned for the optimisation loop; configurations made of the same building blocks
    # (inversion.py:21-25) the generalised native engine - never a torch-module evaluation
    assert type(skip()) is NativeSkip
    inv = skip(32, 3, num_channels_down=[16, 32, 64, 128, 128, 128], num_channels_up=[16, 32, 64, 128, 128, 128],
               num_channels_skip=[4, 4, 4, 4, 4, 4], filter_size_down=[7, 7, 5, 5, 3, 3], filter_size_up=[7, 7, 5, 5, 3, 3],
               downsample_mode='stride', pad='reflection')
    assert type(inv) is NativeSkipX
    assert type(skip(pad="reflection")) is NativeSkipX and type(skip(need_sigmoid=False)) is NativeSkipX
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        inv(torch.zeros(1, 32, 16, 16))
    inv.eval()
    with pytest.raises(NotImplementedError, match="training mode"):
        inv(torch.zeros(1, 32, 16, 16))


def test_scheduler_and_optimizer_factories():
    from splice_b200.util.util import get_optimizer, get_scheduler

    p = [torch.nn.Parameter(torch.zeros(3))]
    cfg = {"optimizer": "sgd", "lr": 0.1}
    opt = get_optimizer(cfg, p)
    assert isinstance(opt, torch.optim.SGD)
    sch = get_scheduler(opt, "none")
    opt.step(); sch.step()
    assert opt.param_groups[0]["lr"] == 0.1
    assert isinstance(get_scheduler(opt, "bogus"), NotImplementedError)          # returned, not raised (ref util.py:24)
    assert isinstance(get_optimizer({"optimizer": "bogus", "lr": 1}, p), NotImplementedError)
    lin = get_scheduler(torch.optim.SGD(p, lr=1.0), "linear", n_epochs_decay=8)
    assert lin.get_last_lr() == [1.0]


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from splice_b200.models.extractor import VitExtractor
    from splice_b200.optim import FusedAdam

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VitExtractor("dino_vits16", "cpu", state_dict={})
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdam([p], lr=1e-3).step()


def test_lambda_schedule_matches_oracle():
    """LossG.update_lambda_config vs the oracle restatement (no engine needed: build the object bare)."""
    from oracle import splice_ref as R
    from splice_b200.util.losses import LossG
    import yaml
    from pathlib import Path

    cfg = yaml.safe_load(open(Path(__file__).resolve().parents[1] / "splice_b200" / "conf" / "default" / "config.yaml"))
    crit = LossG.__new__(LossG)
    crit.cfg = cfg
    crit.lambdas = dict(lambda_global_cls=cfg['lambda_global_cls'], lambda_global_ssim=0, lambda_entire_ssim=0,
                        lambda_entire_cls=0, lambda_global_identity=0)
    state = None
    for step in list(range(0, 5)) + [74, 75, 76, 150, 151]:
        crit.update_lambda_config(torch.tensor([float(step)]))
        state = R.active_lambdas(cfg, step, state)
        assert {k: float(v) for k, v in crit.lambdas.items()} == {k: float(v) for k, v in state.items()}, step


def test_install_as_reference_modules_aliases_every_mirror():
    import subprocess, sys
    from pathlib import Path

    code = ("import splice_b200, sys; splice_b200.install_as_reference_modules();"
            "import models.model, util.losses, util.util, data.Dataset, models.unet.skip, models.extractor;"
            "assert models.model.Model.__module__ == 'splice_b200.models.model';"
            "assert util.losses.LossG.__module__ == 'splice_b200.util.losses';"
            "from models.extractor import VitExtractor, attn_cosine_sim; print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(Path(__file__).resolve().parents[1]))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_generator_slot_policy_over_the_reference_step_schedule():
    """Replays the loop's netG call pattern (models/model.py:12-25 + which outputs each step's loss consumes,
    util/losses.py:34-72): step 0 never backwards y_global, steps = 0 mod 75 keep three passes alive. No pass of
    the current step may be overwritten before its backward."""
    from splice_b200.generator import pick_keep_slot

    tokens, tok = [None] * 4, 0
    for step in range(0, 160):
        entire = step % 75 == 0
        calls = ["x_global"] + (["x_entire"] if entire else []) + ["y_global"]
        used = {"x_global", "x_entire"} if step == 0 else set(calls)
        live = {}
        for name in calls:
            slot = pick_keep_slot(tokens)
            assert slot not in live.values(), (step, name, slot)
            tok += 1
            tokens[slot] = tok
            live[name] = slot
        for name in calls:
            if name in used:
                tokens[live[name]] = None
    # steady state: the call site -> slot mapping is stable (few CUDA-graph keys)
    assert pick_keep_slot(tokens) == 0


def test_prefetched_samples_reproduce_the_inline_stream(tmp_path):
    """The worker-thread feed hands out exactly the samples the reference loop would draw inline:
    same numpy / torch RNG consumption order, `step` snapshotted per sample, the every-75th-step 'A' entry in place."""
    import random

    import numpy as np
    import yaml
    from PIL import Image

    from splice_b200.data.Dataset import SingleImageDataset
    from splice_b200.data.prefetch import PrefetchedSamples

    rng = np.random.default_rng(7)
    for sub in ("A", "B"):
        (tmp_path / sub).mkdir()
        Image.fromarray(rng.integers(0, 256, (96, 96, 3), dtype=np.uint8)).save(tmp_path / sub / "im.png")
    from pathlib import Path
    cfg = yaml.safe_load(open(Path(__file__).resolve().parents[1] / "splice_b200" / "conf" / "default" / "config.yaml"))
    cfg["dataroot"] = str(tmp_path)

    def seed():
        random.seed(3); np.random.seed(3); torch.manual_seed(3)

    n = 80
    seed()
    ds = SingleImageDataset(cfg)
    inline = []
    for _ in range(n):
        s = ds[0]
        inline.append({k: v.clone() for k, v in s.items()})
    for mode in ("thread", "process"):      # the forked worker inherits the generator states the inline loop would draw from
        seed()
        feed = PrefetchedSamples(SingleImageDataset(cfg), n, depth=3, pin=False, mode=mode)
        got = list(feed)
        feed.close()
        assert len(got) == n
        _same_stream(inline, got)


def _same_stream(inline, got):
    for a, b in zip(inline, got):
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), k
    assert "A" in got[0] and "A" in got[75] and "A" not in got[1]
    assert [float(g["step"]) for g in got] == [float(i) for i in range(len(got))]


def test_device_aug_feed_consumes_the_reference_random_stream(tmp_path):
    """data/device_aug.py: same random PARAMETERS as the PIL pipeline (generator states identical after every sample,
    identical crop shapes), pixel arithmetic within the documented bounds of the PIL arithmetic (run on the CPU here:
    the feed only needs a torch device)."""
    import numpy as np
    import torch
    from PIL import Image

    from bench import make_cfg, synth_image
    from splice_b200.data.Dataset import SingleImageDataset
    from splice_b200.data.device_aug import DeviceAugmentedDataset

    for sub, seed, grid in (("A", 1000, 8), ("B", 1001, 16)):
        (tmp_path / sub).mkdir()
        t = synth_image(seed, 200, grid)[:, :150, :200]                      # non-square, like the shipped pairs
        Image.fromarray((t.permute(1, 2, 0).numpy() * 255).round().astype(np.uint8)).save(tmp_path / sub / "im.png")
    cfg = make_cfg("dino_vitb8")
    cfg.update(dataroot=str(tmp_path), global_A_crops_n_crops=2, global_B_crops_n_crops=2, entire_A_every=7)

    def run(make):
        np.random.seed(5)
        torch.manual_seed(5)
        ds = make()
        out = []
        for _ in range(40):
            s = ds[0]
            out.append(({k: v.clone().cpu() for k, v in s.items()}, torch.get_rng_state().clone(), np.random.get_state()[1].copy()))
        return out

    ref = run(lambda: SingleImageDataset(cfg))
    got = run(lambda: DeviceAugmentedDataset(cfg, device="cpu"))
    worst = mean = 0.0
    for (sa, ta, na), (sb, tb, nb) in zip(ref, got):
        assert torch.equal(ta, tb) and (na == nb).all()          # the same draws were made, in the same order
        assert sa.keys() == sb.keys()
        for k in sa:
            assert sa[k].shape == sb[k].shape, k
            d = (sa[k].float() - sb[k].float()).abs()
            worst, mean = max(worst, d.max().item()), max(mean, d.mean().item())
    assert torch.equal(ref[-1][0]["step"], got[-1][0]["step"])
    assert any("A" in s for s, _, _ in got)
    assert mean < 3 / 255 and worst < 24 / 255, (mean * 255, worst * 255)
    # the texture image is only flipped and cropped: exact
    for (sa, _, _), (sb, _, _) in zip(ref, got):
        assert torch.equal(sa["B_global"], sb["B_global"])


def test_inversion_noise_feed_keeps_the_reference_stream():
    """inversion.py:56-62 draws `torch.randn(shape)` inline every iteration; NoiseFeed draws the same tensors ahead on a worker
    thread (same process-wide CPU generator, same order)."""
    from splice_b200.inversion import NoiseFeed

    shape = (1, 32, 20, 30)
    torch.manual_seed(5)
    want = [torch.randn(shape) for _ in range(6)]
    torch.manual_seed(5)
    feed = NoiseFeed(shape, 6, depth=2, pin=False)
    got = [feed.next() for _ in range(6)]
    feed.close()
    assert all(torch.equal(a, b) for a, b in zip(want, got))


def test_inversion_cli_matches_the_reference_script():
    """Same options, types and defaults as the reference's inversion.py:77-92 (checked against its parser in the build
    container; the expected values are spelled out here because /root/reference does not travel)."""
    from splice_b200.inversion import NET_ARGS, build_parser

    got = vars(build_parser().parse_args(["--save_path", "o.png", "--feature", "keys"]))
    assert got == {"feature": "keys", "layer": 11, "dino_model_name": "dino_vitb8",
                   "image_path": "datasets/feature_visualization/limes.jpeg", "save_path": "o.png", "log_freq": 100,
                   "input_depth": 32, "LR": 0.01, "n_iter": 20000, "reduce_noise_stage_1_iter": 10000,
                   "reduce_noise_stage_2_iter": 15000}
    with pytest.raises(SystemExit):
        build_parser().parse_args(["--feature", "cls"])          # --save_path is required
    assert NET_ARGS["filter_size_down"] == [7, 7, 5, 5, 3, 3] and NET_ARGS["pad"] == "reflection"   # inversion.py:21-25


@pytest.mark.skipif(not Path("/root/reference").exists(), reason="needs the reference checkout (build container only)")
def test_reference_scripts_import_over_the_module_swap():
    """with install_as_reference_modules() the reference's OWN train.py and inversion.py import and bind the
    splice_b200 classes (every name they import exists in the mirrors). Running them needs a GPU; importing does not."""
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(1, '/root/reference')\n"
        "import splice_b200; splice_b200.install_as_reference_modules()\n"
        "import train, inversion\n"
        "import splice_b200.models.model as M, splice_b200.util.losses as L, splice_b200.models.extractor as E\n"
        "import splice_b200.models.unet.skip as S, splice_b200.data.Dataset as D\n"
        "assert train.Model is M.Model and train.LossG is L.LossG and train.SingleImageDataset is D.SingleImageDataset\n"
        "assert inversion.VitExtractor is E.VitExtractor and inversion.skip is S.skip\n"
        "assert train.__file__.startswith('/root/reference') and inversion.__file__.startswith('/root/reference')\n"
        "print('ok')\n" % str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-1500:]
