"""CPU interpreter for the engine's static launch list -- TEST INFRASTRUCTURE, never imported by the product.

The engine's host side (siammot_b200/engine.py) turns a model into a list of C-ABI calls over preallocated NHWC buffers:
which buffer feeds which layer, channel-slice views for the concat-free DLA roots, residual operands, strides, the FPN
top-down order.  None of that can be executed in a container without a GPU -- but it is plain data: a list of
(entry point, arguments).  This module builds the plan with the buffers in HOST memory (fp32) and interprets the calls of the
backbone + FPN part with torch CPU ops, reading and writing the very pointers the descriptors carry.  Comparing the FPN maps
with the oracle then checks the wiring itself (tests/test_plan_emulation_cpu.py); the kernels are checked on the GPU.

Interpreted entry points: smot_image_to_nhwc, smot_conv2d, smot_maxpool2x2, smot_maxpool3x3s2, smot_subsample2,
smot_upsample_add, smot_deform_im2col3x3.  Interpretation stops at the first call outside that set (the RPN selection)."""
import ctypes as C

import torch
import torch.nn.functional as F


def _view(ptr, B, H, W, Cc, ld):
    """Writable fp32 NHWC view (B,H,W,Cc) with pixel pitch ld over host memory at ptr."""
    n = (B * H * W - 1) * ld + Cc
    flat = torch.frombuffer((C.c_float * n).from_address(ptr), dtype=torch.float32)
    return flat.as_strided((B, H, W, Cc), (H * W * ld, W * ld, ld, 1))


def _p(a):
    return a.value if isinstance(a, C.c_void_p) else a


def build_engine_on_host(cfg, sd, monkeypatch):
    """An Engine whose buffers live in host memory (construction only: nothing can be launched)."""
    from siammot_b200 import engine as eng_mod
    from siammot_b200 import ops
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(ops, "_require_cuda", lambda *a: None)
    assert cfg.DTYPE == "float32"
    eng = eng_mod.Engine(cfg, device="cpu", use_graph=False)
    eng.load_state_dict(sd)
    return eng


def run_backbone(plan, image):
    """Interpret plan.steps up to the FPN outputs.  image: (3,H,W) fp32.  Returns the number of calls interpreted."""
    plan.img_in.copy_(image)
    done = 0
    for fn, args, tag, _branch in plan.steps:
        if fn in ("fork", "join"):
            continue
        name = getattr(fn, "__name__", None)
        if name == "smot_image_to_nhwc":
            chw, out, Cc, H, W, ld, dt = [_p(a) for a in args]
            assert dt == 0
            src = torch.frombuffer((C.c_float * (Cc * H * W)).from_address(chw), dtype=torch.float32).view(Cc, H, W)
            _view(out, 1, H, W, Cc, ld).copy_(src.permute(1, 2, 0)[None])
        elif name == "smot_conv2d":
            d = args[0]._obj
            assert d.in_dtype == 0 and d.out_dtype == 0
            x = _view(d.inp, d.batch, d.H, d.W, d.Cin, d.in_ld).permute(0, 3, 1, 2)
            nw = d.Cout * d.KH * d.KW * d.Cin
            w = torch.frombuffer((C.c_float * nw).from_address(d.weight), dtype=torch.float32).view(d.Cout, d.KH, d.KW, d.Cin)
            y = F.conv2d(x, w.permute(0, 3, 1, 2), None, d.stride, d.pad)
            assert tuple(y.shape) == (d.batch, d.Cout, d.OH, d.OW), (tag, tuple(y.shape), (d.OH, d.OW))
            if d.scale:
                y = y * torch.frombuffer((C.c_float * d.Cout).from_address(d.scale), dtype=torch.float32).view(1, -1, 1, 1)
            if d.bias:
                y = y + torch.frombuffer((C.c_float * d.Cout).from_address(d.bias), dtype=torch.float32).view(1, -1, 1, 1)
            if d.residual:
                y = y + _view(d.residual, d.batch, d.OH, d.OW, d.Cout, d.res_ld).permute(0, 3, 1, 2)
            if d.relu:
                y = F.relu(y)
            _view(d.out, d.batch, d.OH, d.OW, d.Cout, d.out_ld).copy_(y.permute(0, 2, 3, 1))
        elif name in ("smot_maxpool2x2", "smot_maxpool3x3s2"):
            inp, out, B, H, W, Cc, ild, old, dt = [_p(a) for a in args]
            x = _view(inp, B, H, W, Cc, ild).permute(0, 3, 1, 2)
            y = F.max_pool2d(x, 2, 2) if name == "smot_maxpool2x2" else F.max_pool2d(x, 3, 2, 1)
            _view(out, B, y.shape[2], y.shape[3], Cc, old).copy_(y.permute(0, 2, 3, 1))
        elif name == "smot_subsample2":
            inp, out, H, W, Cc, ild, old, dt = [_p(a) for a in args]
            y = _view(inp, 1, H, W, Cc, ild)[:, ::2, ::2]
            _view(out, 1, y.shape[1], y.shape[2], Cc, old).copy_(y)
        elif name == "smot_deform_im2col3x3":
            from cabi_emulator import deform_columns
            inp, off, cols, H, W, Cc, ild, oild, OH, OW, old, stride, dt = [_p(a) for a in args]
            deform_columns(inp, off, cols, H, W, Cc, ild, oild, OH, OW, old, stride)
        elif name == "smot_upsample_add":
            top, Ht, Wt, tld, lat, H, W, lld, Cc, dt = [_p(a) for a in args]
            t = _view(top, 1, Ht, Wt, Cc, tld).permute(0, 3, 1, 2)
            up = F.interpolate(t, size=(H, W), mode="bilinear", align_corners=False)
            lv = _view(lat, 1, H, W, Cc, lld)
            lv.copy_(lv + up.permute(0, 2, 3, 1))
        else:
            break
        done += 1
    return done
