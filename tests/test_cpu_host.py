"""CPU tests (no GPU needed): the C-ABI library loads and exports every symbol the header declares, host
logic (schedules, mask generator, flat layout rules, bucketed all-reduce under gloo with world_size 2)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "vjepa2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vj_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from vjepa2_b200 import _cabi
    lib = _cabi.load()                       # raises if the .so is missing
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vjepa2_b200.h but not exported"
        assert s in _cabi.SIGNATURES, f"{s} has no ctypes prototype"
    assert set(_cabi.SIGNATURES) == set(syms)
    assert lib.vj_abi_version() == 3


def test_product_path_has_no_cpu_fallback():
    """A CPU tensor must raise, not silently run something else; and nothing in the package imports the oracle."""
    from vjepa2_b200.masks import apply_masks
    with pytest.raises(RuntimeError):
        apply_masks(torch.zeros(1, 8, 8), [torch.zeros(1, 2, dtype=torch.int64)])
    pkg = os.path.join(ROOT, "vjepa2_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "vjepa_oracle" not in src and "import oracle" not in src, fn
    from functools import partial
    import torch.nn as nn
    from vjepa2_b200.vision_transformer import VisionTransformer
    m = VisionTransformer(img_size=32, patch_size=16, num_frames=4, embed_dim=64, depth=1, num_heads=1,
                          norm_layer=partial(nn.LayerNorm, eps=1e-6), use_rope=True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 4, 32, 32))     # parameters on CPU -> loud failure


def test_module_tree_and_param_names_match_reference_layout():
    from functools import partial
    import torch.nn as nn
    import vjepa_oracle as O
    from vjepa2_b200.predictor import vit_predictor
    from vjepa2_b200.vision_transformer import VisionTransformer
    enc = VisionTransformer(img_size=96, patch_size=16, num_frames=8, embed_dim=128, depth=2, num_heads=2,
                            norm_layer=partial(nn.LayerNorm, eps=1e-6), use_rope=True)
    w = O.init_encoder_weights(128, 2, 4.0)
    assert list(enc.state_dict().keys()) == list(w.keys())
    assert all(enc.state_dict()[k].shape == v.shape for k, v in w.items())
    pred = vit_predictor(img_size=96, patch_size=16, num_frames=8, embed_dim=128, predictor_embed_dim=64, depth=2,
                         num_heads=2, use_mask_tokens=True, num_mask_tokens=2, use_rope=True)
    wp = O.init_predictor_weights(128, 64, 2, 2)
    assert list(pred.state_dict().keys()) == list(wp.keys())
    # reference init statistics (vision_transformer.py:130-153): LN = (1, 0), biases 0, proj/fc2 rescaled
    assert float(enc.blocks[1].attn.proj.weight.std()) < float(enc.blocks[1].attn.qkv.weight.std())
    assert float(pred.mask_tokens[0].abs().max()) == 0.0
    for name in ("vit_large", "vit_giant_xformers"):
        from vjepa2_b200 import vision_transformer as V
        assert callable(getattr(V, name))


def test_schedules_match_reference(golden):
    from vjepa2_b200.schedulers import CosineWDSchedule, WarmupCosineSchedule, momentum_schedule
    s = WarmupCosineSchedule(warmup_steps=4, start_lr=1e-4, ref_lr=5.25e-4, T_max=20, final_lr=1e-5)
    w = CosineWDSchedule(ref_wd=0.04, T_max=20, final_wd=0.4)
    lr = torch.tensor([s.step() for _ in range(24)], dtype=torch.float64)
    wd = torch.tensor([w.step() for _ in range(24)], dtype=torch.float64)
    torch.testing.assert_close(lr, golden["sched.lr"], rtol=1e-12, atol=0)
    torch.testing.assert_close(wd, golden["sched.wd"], rtol=1e-12, atol=0)
    m = list(momentum_schedule((0.99, 1.0), 10, 2, 1.25))
    assert len(m) == 26 and m[0] == 0.99 and abs(m[-1] - 1.0) < 1e-12


def test_mask_collator_bit_exact(golden):
    """Same torch CPU RNG state in -> bit-identical indices out."""
    import vjepa_oracle as O
    from vjepa2_b200.masks import MaskCollator
    cfgs = [dict(m) for m in O.DEFAULT_MASK_CFG]
    coll = MaskCollator(cfgs_mask=cfgs, dataset_fpcs=[16], crop_size=(256, 256), patch_size=(16, 16), tubelet_size=2)
    torch.manual_seed(239)
    for it in range(3):
        batch = [(torch.zeros(1), 0, [list(range(16))]) for _ in range(6)]
        _, enc, pred = coll(batch)[0]
        for j in range(2):
            assert enc[j].dtype == torch.int64
            assert torch.equal(enc[j], golden[f"mask.it{it}.enc{j}"])
            assert torch.equal(pred[j], golden[f"mask.it{it}.pred{j}"])
            # structural invariants: ascending, disjoint, in range
            assert bool((enc[j][:, 1:] > enc[j][:, :-1]).all()) and bool((pred[j][:, 1:] > pred[j][:, :-1]).all())
            for b in range(6):
                assert not set(enc[j][b].tolist()) & set(pred[j][b].tolist())
            assert int(pred[j].max()) < 2048 and enc[j].shape[1] % 8 == 0
    coll2 = MaskCollator(cfgs_mask=cfgs, dataset_fpcs=[64], crop_size=(384, 384), patch_size=(16, 16), tubelet_size=2)
    torch.manual_seed(7)
    enc, pred = coll2.draw(64, 2)
    for j in range(2):
        assert torch.equal(enc[j], golden[f"mask384.enc{j}"]) and torch.equal(pred[j], golden[f"mask384.pred{j}"])


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vjepa2_b200.train import GradBucketer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
g = torch.Generator().manual_seed(100 + rank)          # rank-local "gradients" (bench: seed = base + rank)
flat = torch.randn(10 * 1024, generator=g)
mine = flat.clone()
b = GradBucketer()
assert b.world == world
for start in range(9 * 1024, -1, -1024 * 3):          # reverse order, like backward
    b.submit(flat, start, min(start + 3 * 1024, flat.numel()))
b.wait()
ref = sum(torch.randn(10 * 1024, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
assert torch.allclose(flat, ref, atol=1e-6), float((flat - ref).abs().max())
# data-parallel averaging is folded into inv_scale = 1/(scale*world)
avg = flat * (1.0 / world)
assert torch.allclose(avg, ref / world)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_bucketed_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()
        assert b"ok" in out


# ---------------------------------------------------------------------------------------------------------------
# checkpoint wire format (train.py:315-333, app/vjepa/utils.py:90-135, src/hub/backbones.py:22-28)
# ---------------------------------------------------------------------------------------------------------------
def _tiny_models():
    from functools import partial
    import torch.nn as nn
    from vjepa2_b200.predictor import vit_predictor
    from vjepa2_b200.vision_transformer import VisionTransformer
    from vjepa2_b200.wrappers import MultiSeqWrapper, PredictorMultiSeqWrapper
    enc = VisionTransformer(img_size=32, patch_size=16, num_frames=4, embed_dim=64, depth=2, num_heads=2,
                            norm_layer=partial(nn.LayerNorm, eps=1e-6), use_rope=True)
    pred = vit_predictor(img_size=32, patch_size=16, num_frames=4, embed_dim=64, predictor_embed_dim=64, depth=1,
                         num_heads=2, use_mask_tokens=True, num_mask_tokens=2, zero_init_mask_tokens=True,
                         use_rope=True)
    return MultiSeqWrapper(enc), PredictorMultiSeqWrapper(pred)


def _reference_style_optimizer(encoder, predictor):
    # init_opt's four groups, written against the WRAPPED modules exactly as app/vjepa/utils.py:224-239 does
    groups = [
        {"params": (p for n, p in encoder.named_parameters() if ("bias" not in n) and (len(p.shape) != 1))},
        {"params": (p for n, p in predictor.named_parameters() if ("bias" not in n) and (len(p.shape) != 1))},
        {"params": (p for n, p in encoder.named_parameters() if ("bias" in n) or (len(p.shape) == 1)),
         "WD_exclude": True, "weight_decay": 0},
        {"params": (p for n, p in predictor.named_parameters() if ("bias" in n) or (len(p.shape) == 1)),
         "WD_exclude": True, "weight_decay": 0},
    ]
    return torch.optim.AdamW(groups, betas=(0.9, 0.999), eps=1e-8)


def test_optimizer_state_dict_matches_torch_adamw_layout():
    from vjepa2_b200 import checkpoint as C
    enc, pred = _tiny_models()
    opt = _reference_style_optimizer(enc, pred)
    g = torch.Generator().manual_seed(3)
    unused = pred.backbone.mask_tokens[1]
    for p in list(enc.parameters()) + list(pred.parameters()):
        if p is not unused:                           # a parameter without a gradient gets no optimizer state
            p.grad = torch.randn(p.shape, generator=g)
    for grp in opt.param_groups:
        grp["lr"] = 1e-3
        if not grp.get("WD_exclude", False):
            grp["weight_decay"] = 0.04
    opt.step()
    opt.step()
    ref_sd = opt.state_dict()

    groups = C.opt_param_groups(enc, pred)
    by_param = {}
    for grp in opt.param_groups:
        for p in grp["params"]:
            if p in opt.state:
                by_param[id(p)] = (opt.state[p]["exp_avg"], opt.state[p]["exp_avg_sq"])
    ours = C.build_opt_state_dict(groups, lambda t: by_param.get(id(t)), step=2, lr=1e-3, wd=0.04)

    assert [g_["params"] for g_ in ours["param_groups"]] == [g_["params"] for g_ in ref_sd["param_groups"]]
    for a, b in zip(ours["param_groups"], ref_sd["param_groups"]):
        assert set(a.keys()) == set(b.keys())
        for k in ("lr", "weight_decay", "betas", "eps", "amsgrad"):
            assert a[k] == b[k], k
        assert a.get("WD_exclude", None) == b.get("WD_exclude", None)
    assert set(ours["state"].keys()) == set(ref_sd["state"].keys())
    for k, s in ref_sd["state"].items():
        assert float(ours["state"][k]["step"]) == float(s["step"]) == 2.0
        assert torch.equal(ours["state"][k]["exp_avg"], s["exp_avg"])
        assert torch.equal(ours["state"][k]["exp_avg_sq"], s["exp_avg_sq"])

    # torch's own loader accepts it, and the inverse mapping restores every moment bit for bit
    opt2 = _reference_style_optimizer(enc, pred)
    opt2.load_state_dict(ours)
    bufs = {id(t): (torch.full(t.shape, 7.0), torch.full(t.shape, 7.0)) for grp in groups for _, t in grp}
    step = C.restore_opt_state(groups, ref_sd, lambda t: bufs[id(t)])
    assert step == 2
    for grp in groups:
        for _, t in grp:
            if t is unused:
                assert float(bufs[id(t)][0].abs().sum()) == 0.0
            else:
                assert torch.equal(bufs[id(t)][0], by_param[id(t)][0])
                assert torch.equal(bufs[id(t)][1], by_param[id(t)][1])


def test_checkpoint_key_prefixes_and_pretrained_loading():
    from vjepa2_b200 import checkpoint as C
    enc, _ = _tiny_models()
    sd = {"module.backbone." + k: v.clone() + 1.0 for k, v in enc.backbone.state_dict().items()}
    sd["module.backbone.pos_embed"] = torch.zeros(1, 8, 64)               # sincos checkpoints carry one (backbones.py:132)
    cleaned = C.clean_backbone_key(sd)
    assert "patch_embed.proj.weight" in cleaned and not any(k.startswith("module") for k in cleaned)
    bad = dict(sd)
    bad["module.backbone.norm.weight"] = torch.zeros(3)                   # wrong shape: the model's tensor is kept
    before = enc.backbone.norm.weight.detach().clone()
    msg = C.load_pretrained(enc, {"target_encoder": bad}, checkpoint_key="target_encoder")
    assert list(msg.unexpected_keys) == ["pos_embed"] and not msg.missing_keys
    assert torch.equal(enc.backbone.norm.weight, before)
    assert torch.equal(enc.backbone.blocks[0].attn.qkv.weight, sd["module.backbone.blocks.0.attn.qkv.weight"])
    assert C._prefix_of(enc) == "backbone."

    class _DDP(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.module = m
    assert C._prefix_of(_DDP(enc)) == "module.backbone."


def test_clip_aggregation_regroups_views_and_clips_like_the_reference():
    """vit_encoder_multiclip.py:107-149 with a stand-in encoder (token = clip statistics), so the regrouping of
    [clips x views x batch] into per-view time-concatenated token lists is checked without a GPU."""
    from vjepa2_b200.inference import ClipAggregation

    class _Enc(torch.nn.Module):
        embed_dim, num_heads, tubelet_size = 4, 1, 2

        def forward(self, x):                                   # [n, C, F, H, W] -> [n, (F/2)*S, D], S = 3
            n, _, F, _, _ = x.shape
            t = x.reshape(n, 3, F // 2, 2, -1).mean(dim=(1, 3, 4))           # [n, T]
            return t[:, :, None, None].expand(n, F // 2, 3, 4).reshape(n, -1, 4).contiguous()

    B, F = 2, 4
    g = torch.Generator().manual_seed(0)
    x = [[torch.randn(B, 3, F, 8, 8, generator=g) for _ in range(3)] for _ in range(2)]     # 2 clips x 3 views
    agg = ClipAggregation(_Enc(), tubelet_size=2)
    outs = agg(x)
    assert len(outs) == 3 and outs[0].shape == (B, 2 * (F // 2) * 3, 4)
    enc = _Enc()
    for j in range(3):
        want = torch.cat([enc(x[i][j]).reshape(B, F // 2, 3, 4) for i in range(2)], dim=1).flatten(1, 2)
        assert torch.equal(outs[j], want)


def test_hub_constructors_build_reference_shapes():
    from vjepa2_b200 import inference as I
    enc, pred = I.vjepa2_vit_large(num_frames=16)
    assert enc.embed_dim == 1024 and len(enc.blocks) == 24 and enc.num_patches == 8 * 16 * 16
    assert pred.predictor_embed.weight.shape == (384, 1024) and len(pred.mask_tokens) == 10
    with pytest.raises(RuntimeError):
        I.vjepa2_vit_large(pretrained=True)                    # no download path: a local checkpoint is required
    ck = {"encoder": {"module.backbone." + k: v for k, v in enc.state_dict().items()},
          "predictor": {"module.backbone." + k: v for k, v in pred.state_dict().items()}}
    ck["encoder"]["module.backbone.pos_embed"] = torch.zeros(1, 2048, 1024)
    enc2, pred2 = I.vjepa2_vit_large(pretrained=True, checkpoint=ck, num_frames=16)
    assert torch.equal(enc2.blocks[3].mlp.fc1.weight, enc.blocks[3].mlp.fc1.weight)
    assert torch.equal(pred2.predictor_proj.weight, pred.predictor_proj.weight)


def test_mask_passes_respect_the_activation_budget():
    """JepaTrainStep runs all masks of a group in one pass unless their saved activations exceed the budget
    (64f x 384px geometry); then it goes pass by pass, never splitting a mask and keeping the order."""
    from types import SimpleNamespace
    from vjepa2_b200.train import JepaTrainStep
    enc, pred = _tiny_models()
    B = 4
    mes = [torch.zeros(B, 100, dtype=torch.int64), torch.zeros(B, 30, dtype=torch.int64), torch.zeros(B, 60, dtype=torch.int64)]
    mps = [torch.zeros(B, 50, dtype=torch.int64), torch.zeros(B, 120, dtype=torch.int64), torch.zeros(B, 90, dtype=torch.int64)]
    ns = SimpleNamespace(encoder=enc.backbone, predictor=pred.backbone, ACT_BUDGET_BYTES=JepaTrainStep.ACT_BUDGET_BYTES)
    assert JepaTrainStep._mask_passes(ns, mes, mps) == [(0, 3)]
    per_enc = 2 * (16 * 64 + 4 * 256)
    per_pred = 1 * (20 * 64 + 4 * 256)
    need = [int(1.1 * (me.numel() * per_enc + (me.numel() + mp.numel()) * per_pred)) for me, mp in zip(mes, mps)]
    ns.ACT_BUDGET_BYTES = need[0] + need[1]              # first two fit together, the third starts a new pass
    assert JepaTrainStep._mask_passes(ns, mes, mps) == [(0, 2), (2, 3)]
    ns.ACT_BUDGET_BYTES = 1                               # nothing fits: one mask per pass, never an empty pass
    assert JepaTrainStep._mask_passes(ns, mes, mps) == [(0, 1), (1, 2), (2, 3)]


def test_wrappers_fan_out_protocol():
    """src/utils/wrappers.py:15-43: outputs nest as [group][mask]; the predictor gets mask_index = group index."""
    from vjepa2_b200.wrappers import MultiSeqWrapper, PredictorMultiSeqWrapper
    calls = []

    class _Enc(torch.nn.Module):
        def forward(self, x, masks=None):
            calls.append(("enc", int(x), None if masks is None else int(masks)))
            return (int(x), None if masks is None else int(masks))

    class _Pred(torch.nn.Module):
        def forward(self, z, mx, my, mask_index=1, has_cls=False):
            return (z, int(mx), int(my), mask_index, has_cls)

    enc = MultiSeqWrapper(_Enc())
    x = [torch.tensor(10), torch.tensor(20)]
    masks = [[torch.tensor(1), torch.tensor(2)], [torch.tensor(3)]]
    assert enc(x) == [(10, None), (20, None)]
    z = enc(x, masks)
    assert z == [[(10, 1), (10, 2)], [(20, 3)]]
    assert [c[1:] for c in calls[-3:]] == [(10, 1), (10, 2), (20, 3)]          # group-major call order
    pred = PredictorMultiSeqWrapper(_Pred())
    out = pred(z, masks, [[torch.tensor(7), torch.tensor(8)], [torch.tensor(9)]], has_cls=False)
    assert out == [[((10, 1), 1, 7, 0, False), ((10, 2), 2, 8, 0, False)], [((20, 3), 3, 9, 1, False)]]
    assert list(dict(enc.named_modules()).keys())[1] == "backbone"             # checkpoint key prefix


# ---------------------------------------------------------------------------------------------------------------
# boundary hygiene: the ctypes struct must follow the header field for field
# ---------------------------------------------------------------------------------------------------------------
def _header_struct_fields(name):
    txt = open(os.path.join(ROOT, "include", "vjepa2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    end = re.search(r"\}\s*" + name + r"\s*;", txt).start()
    start = [m.end() for m in re.finditer(r"typedef\s+struct\s*\{", txt[:end])][-1]     # the struct that closes there
    body = txt[start:end]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        first, *more = [d.strip() for d in decl.split(",")]
        ctype, fname = first.rsplit(None, 1) if "*" not in first else (first[:first.rindex("*") + 1], first[first.rindex("*") + 1:])
        ctype = " ".join(ctype.split())
        fields.append((fname.strip(), ctype))
        fields += [(m, ctype) for m in more]
    return fields


def test_gemm_args_struct_matches_header_field_for_field():
    import ctypes
    from vjepa2_b200 import _cabi
    want = _header_struct_fields("vj_gemm_args")
    got = _cabi.GemmArgs._fields_
    assert [n for n, _ in want] == [n for n, _ in got]
    for (n, ctype), (_, ct) in zip(want, got):
        if "*" in ctype:
            assert ct is ctypes.c_void_p, n
        elif ctype == "int64_t":
            assert ct is ctypes.c_int64, n
        elif ctype == "int32_t":
            assert ct is ctypes.c_int32, n
        else:
            raise AssertionError(f"unexpected C type {ctype!r} for {n}")
    # INTEGRATION.md shows the same struct to maintainers: every field must appear there, in order
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    pos = [doc.find(f'("{n}"') for n, _ in got]
    assert all(p >= 0 for p in pos) and pos == sorted(pos), "INTEGRATION.md GemmArgs snippet is out of sync"
    assert _cabi.ABI_VERSION == 3


def test_mask_spec_struct_matches_header_field_for_field():
    import ctypes
    from vjepa2_b200 import _cabi
    want = _header_struct_fields("vj_mask_spec")
    got = _cabi.MaskSpec._fields_
    assert [n for n, _ in want] == [n for n, _ in got]
    for (n, ctype), (_, ct) in zip(want, got):
        assert ct is {"int32_t": ctypes.c_int32, "double": ctypes.c_double}[ctype], n
    assert _cabi.MASK_RNG_WORDS == 626 and ctypes.sizeof(_cabi.PtrList) == 8 * _cabi.MAX_PEERS
    hdr = open(os.path.join(ROOT, "include", "vjepa2_b200.h")).read()
    assert "#define VJ_MASK_RNG_WORDS 626" in hdr and f"#define VJ_MAX_PEERS {_cabi.MAX_PEERS}" in hdr


def test_device_mask_collator_rng_state_conversion_round_trips():
    """torch.get_rng_state() <-> the 626 words vj_mask_collate keeps on the device (host-side plumbing, no GPU)."""
    from vjepa2_b200 import masks as M
    torch.manual_seed(5)
    for _ in range(700):                                  # cross a refill of the 624-word state
        torch.randint(0, 9, (1,))
    st = torch.get_rng_state()
    w = M._mt_words_from_torch_state(st)
    assert w.numel() == 626 and w.dtype == torch.int32
    assert torch.equal(M._torch_state_from_mt_words(w, st), st)
    # the words really are the generator: restoring them into a scrambled state reproduces the stream
    want = [int(torch.randint(0, 1000, (1,))) for _ in range(5)]
    torch.manual_seed(99)
    torch.set_rng_state(M._torch_state_from_mt_words(w, torch.get_rng_state()))
    assert [int(torch.randint(0, 1000, (1,))) for _ in range(5)] == want


def test_stale_library_abi_is_rejected(monkeypatch):
    from vjepa2_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "ABI_VERSION", 999)
    with pytest.raises(RuntimeError, match="ABI version"):
        _cabi.load()
    monkeypatch.setattr(_cabi, "ABI_VERSION", 3)
    assert _cabi.load() is not None


def test_default_checkpoint_prefix_strict_loads_into_ddp_wrapped_reference_layout():
    """app/vjepa/utils.py:104-118 strict-loads into DDP(MultiSeqWrapper(model)): keys must be module.backbone.*"""
    import torch.nn as nn
    from vjepa2_b200 import checkpoint as C
    enc, _ = _tiny_models()

    class _DDP(nn.Module):                       # the wrapper nesting of train.py:279 without a process group
        def __init__(self, m):
            super().__init__()
            self.module = m
    sd = C._wrapped_state_dict(enc, C.DEFAULT_PREFIX)
    missing = _DDP(enc).load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert set(C.clean_backbone_key(sd)) == set(enc.backbone.state_dict())


# ---------------------------------------------------------------------------------------------------------------
# find_unused_parameters semantics across ranks with UNEQUAL fpc-group counts (train.py:280)
# ---------------------------------------------------------------------------------------------------------------
_FROZEN_WORKER = r"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vjepa2_b200.train import FrozenTokenSync
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=2)
# 8 tiles; tokens 0..2 live in tiles 2, 3 and 5 (tile 5 carries the weight-decay bit like a real mask token)
flags = torch.zeros(8, dtype=torch.uint8)
flags[[2, 3, 5]] = 1
sync = FrozenTokenSync(flags, [(2, 1), (3, 1), (5, 1)])
# rank 0 sees one fpc group (uses token 0), rank 1 sees two (tokens 0 and 1): token 1 is updated on BOTH ranks
used = sync.update({0} if rank == 0 else {0, 1})
assert used.tolist() == [1, 1, 0], used
assert flags.tolist() == [0, 0, 1, 1, 0, 3, 0, 0], flags.tolist()
# next step: nobody uses token 1 any more, rank 0 alone uses token 2
sync.update({0, 2} if rank == 0 else {0})
assert flags.tolist() == [0, 0, 1, 3, 0, 1, 0, 0], flags.tolist()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_frozen_mask_tokens_agree_across_ranks_gloo_world2(tmp_path):
    script = tmp_path / "frozen_worker.py"
    script.write_text(_FROZEN_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()
        assert b"ok" in out


def test_frozen_token_sync_single_process():
    from vjepa2_b200.train import FrozenTokenSync
    flags = torch.tensor([1, 1, 1, 0], dtype=torch.uint8)
    sync = FrozenTokenSync(flags, [(0, 1), (1, 1), (2, 1)])
    sync.update({0})
    assert flags.tolist() == [1, 3, 3, 0]
    assert sync.update({0}) is None                   # unchanged key: nothing to do
    sync.update({0, 1, 2})
    assert flags.tolist() == [1, 1, 1, 0]


def test_mask_mirror_matches_reference_on_the_full_2048_token_grid():
    """The shipped mask config on the real 8x16x16 grid, config seed 239: indices bit-exact against the reference's
    MaskCollator (stored in the full-width golden file)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fullwidth_common as FW
    from vjepa2_b200.masks import MaskCollator
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "ref_fullwidth.pt"), map_location="cpu")
    me, mp = FW.draw_masks(MaskCollator)
    for j in range(2):
        assert torch.equal(me[j], gold[f"vit_large.masks_enc.{j}"].long())
        assert torch.equal(mp[j], gold[f"vit_large.masks_pred.{j}"].long())
        assert me[j].dtype == torch.int64 and int(me[j].max()) < 2048


def test_bucket_hook_merges_adjacent_block_ranges_and_covers_everything_once():
    from vjepa2_b200.train import make_bucket_hook
    depth, blk = 6, 1000
    ranges = {-1: (0, 300)}
    ranges.update({i: (300 + i * blk, 300 + (i + 1) * blk) for i in range(depth)})
    ranges[depth] = (300 + depth * blk, 300 + depth * blk + 50)
    order = [depth] + list(range(depth - 1, -1, -1)) + [-1]       # engine.encoder_backward's completion order
    for bucket_elems, want_max in ((0, len(order)), (2500, 4), (10 ** 9, 1)):
        got = []
        hook = make_bucket_hook(ranges, bucket_elems * 4, lambda lo, hi: got.append((lo, hi)))
        for k in order:
            hook(k)
        assert len(got) <= want_max
        covered = sorted(got)
        assert covered[0][0] == 0 and covered[-1][1] == ranges[depth][1]
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))       # contiguous, no overlap, no gap
    assert len(got) == 1


def test_bias_grad_pad_rule():
    """LayerNorm pads its saved output with 8 ones-columns only where the wgrad GEMM can carry them for free: the last
    256-column tile of D is partly filled, and every weight-gradient GEMM reading it is large enough for the CTA-pair
    kernel (vjepa2_b200/ops.py: bias_grad_pad)."""
    from vjepa2_b200 import ops
    assert ops.bias_grad_pad(1408, 4224, 6144) == 8          # ViT-g: 1408 = 5 * 256 + 128
    assert ops.bias_grad_pad(384, 1152, 1536) == 8           # predictor
    assert ops.bias_grad_pad(1024, 3072, 4096) == 0          # ViT-L: D fills its tiles, an extra tile would cost more
    assert ops.bias_grad_pad(1280, 3840, 5120) == 0          # ViT-H
    assert ops.bias_grad_pad(128, 384, 512) == 0             # toy widths: wgrads below the pair kernel's M >= 1024
    assert ops.bias_grad_pad(1408, 4224, 512) == 0
