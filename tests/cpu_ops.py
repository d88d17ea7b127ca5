"""TEST INFRASTRUCTURE: plain-PyTorch stand-ins with the exact call contracts of hallo_b200.ops (the ctypes wrappers of
the C ABI), used by tests/test_plan_cpu.py to execute the HOST side of the product -- the kernel plans of
hallo_b200/engine.py and hallo_b200/refnet.py: buffer reuse, row offsets, frame <-> pixel bookkeeping, weight packing,
window hoisting, the step tail -- on a machine without a GPU.  Nothing under hallo_b200/ imports this file; the product
has no CPU path.  Each function restates what include/hallo_b200.h specifies for the entry point it stands in for."""
import math

import torch
import torch.nn.functional as F


def install(monkeypatch):
    from hallo_b200 import ops
    for name, fn in list(globals().items()):
        if name.startswith("op_"):
            monkeypatch.setattr(ops, name[3:], fn)

    class _S:
        def synchronize(self):
            pass

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _S())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)


def _gelu(x):
    return F.gelu(x)


def op_gemm(a, w, out, *, bias=None, residual=None, row_scale=None, group_bias=None, rows_per_group=0, alpha=1.0,
            geglu=False, silu=False, relu=False, a2=None, ln_stats=None, ln_colsum=None, ln_eps=1e-5, stats_out=None,
            scatter=None):
    assert ln_stats is None and stats_out is None and scatter is None
    x = a if a2 is None else torch.cat([a, a2], dim=1)
    v = x.float() @ w.float().t()
    if bias is not None:
        v = v + bias.float()
    if group_bias is not None:
        rows = torch.arange(v.shape[0]) // rows_per_group
        v = v + group_bias.float()[rows]
    if silu:
        v = F.silu(v)
    elif relu:
        v = torch.relu(v)
    if geglu:
        v = v[:, 0::2] * _gelu(v[:, 1::2])
    if row_scale is not None:
        v = v * row_scale.float()[:, None]
    v = v * alpha
    if residual is not None:
        v = v + residual.float()
    out.copy_(v.to(out.dtype))
    return out


def _unpack_conv(w_packed, cin):
    cout = w_packed.shape[0]
    return w_packed.float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)


def op_conv3x3(x, w_packed, out, *, bias=None, residual=None, group_bias=None, rows_per_group=0):
    n, h, w_, cin = x.shape
    y = F.conv2d(x.float().permute(0, 3, 1, 2), _unpack_conv(w_packed, cin), None, padding=1)
    v = y.permute(0, 2, 3, 1).reshape(n * h * w_, -1)
    if bias is not None:
        v = v + bias.float()
    if group_bias is not None:
        v = v + group_bias.float()[torch.arange(v.shape[0]) // rows_per_group]
    if residual is not None:
        v = v + residual.float()
    out.copy_(v[:, :out.shape[1]].to(out.dtype))
    return out


def op_phase_split(x, out):
    n, h, w_, c = x.shape
    o = out.view(4, n, h // 2, w_ // 2, c)
    for p in range(2):
        for q in range(2):
            o[p * 2 + q].copy_(x[:, p::2, q::2])
    return out


def op_conv3x3_stride2(x_planes, w_packed, out, *, n, ho, wo, bias=None):
    cin = x_planes.shape[-1]
    pl = x_planes.view(4, n, ho, wo, cin)
    x = torch.empty(n, 2 * ho, 2 * wo, cin, dtype=x_planes.dtype)
    for p in range(2):
        for q in range(2):
            x[:, p::2, q::2] = pl[p * 2 + q]
    y = F.conv2d(x.float().permute(0, 3, 1, 2), _unpack_conv(w_packed, cin), None, stride=2, padding=1)
    v = y.permute(0, 2, 3, 1).reshape(n * ho * wo, -1)
    if bias is not None:
        v = v + bias.float()
    out.copy_(v.to(out.dtype))
    return out


def op_upsample2x(x, out):
    out.copy_(x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    return out


def _sdpa(q, k, v, heads):
    # q [B, Lq, C], k/v [B, Lk, C]
    B, Lq, C = q.shape
    d = C // heads
    qh = q.float().view(B, Lq, heads, d).transpose(1, 2)
    kh = k.float().view(B, -1, heads, d).transpose(1, 2)
    vh = v.float().view(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, C)


def op_attention(q, k, v, out, *, heads, L, kref=None, vref=None, ref_index=None):
    frames = q.shape[0] // L
    C = q.shape[1]
    q3, k3, v3 = (t.reshape(frames, L, C) for t in (q, k, v))
    res = []
    for n in range(frames):
        kk, vv = k3[n:n + 1], v3[n:n + 1]
        r = -1 if ref_index is None else int(ref_index[n])
        if r >= 0:
            kk = torch.cat([kk, kref.reshape(-1, L, C)[r:r + 1]], 1)
            vv = torch.cat([vv, vref.reshape(-1, L, C)[r:r + 1]], 1)
        res.append(_sdpa(q3[n:n + 1], kk, vv, heads))
    out.copy_(torch.cat(res, 0).reshape(frames * L, C).to(out.dtype))
    return out


def op_layernorm(x, gamma, beta, out, *, eps=1e-5, pe=None, pe_index=None, tokens_per_frame=0, frames=0):
    y = F.layer_norm(x.float(), (x.shape[1],), gamma.float(), beta.float(), eps)
    if pe is not None:
        fr = (torch.arange(x.shape[0]) // tokens_per_frame) % frames
        idx = fr if pe_index is None else pe_index.long()[fr]
        y = y + pe.float()[idx]
    out.copy_(y.to(out.dtype))
    return out


def op_groupnorm(x1, gamma, beta, out, stats_ws, *, n_frames, hw, groups=32, eps=1e-5, silu=False, x2=None, fpb_in=0,
                 fpb_out=0, frame_off=0):
    x = x1 if x2 is None else torch.cat([x1, x2], dim=1)
    C = x.shape[1]
    y = F.group_norm(x.float().view(n_frames, hw, C).permute(0, 2, 1), groups, gamma.float(), beta.float(), eps)
    if silu:
        y = F.silu(y)
    y = y.permute(0, 2, 1)
    if fpb_in <= 0:
        fpb_in, fpb_out, frame_off = n_frames, n_frames, 0
    o = out.view(-1, hw, C)
    for n in range(n_frames):
        o[(n // fpb_in) * fpb_out + frame_off + n % fpb_in].copy_(y[n].to(out.dtype))
    return out


def op_cross_attention(q, k, v, out, *, frames, tokens, heads, head_dim, n_keys, kv_frame_div=1, regions=1,
                       q_region_stride=0, kv_region_stride=0, o_region_stride=0):
    C = heads * head_dim

    def region(t, off):          # the kernel does pointer arithmetic: region r starts `off` elements after the view's base
        return t.as_strided((t.shape[0], C), t.stride(), t.storage_offset() + off)

    for r in range(regions):
        qr = region(q, r * q_region_stride).reshape(frames, tokens, C)
        kr = region(k, r * kv_region_stride).reshape(-1, n_keys, C)
        vr = region(v, r * kv_region_stride).reshape(-1, n_keys, C)
        idx = torch.arange(frames) // kv_frame_div
        o = _sdpa(qr, kr[idx], vr[idx], heads)
        region(out, r * o_region_stride).copy_(o.reshape(frames * tokens, C).to(out.dtype))
    return out


def op_temporal_attention(q, k, v, out, *, batch, fq, fk, tokens, heads):
    C = out.shape[1]
    qq = q.reshape(batch, fq, tokens, C).permute(0, 2, 1, 3).reshape(batch * tokens, fq, C)
    kk = k.reshape(batch, fk, tokens, C).permute(0, 2, 1, 3).reshape(batch * tokens, fk, C)
    vv = v.reshape(batch, fk, tokens, C).permute(0, 2, 1, 3).reshape(batch * tokens, fk, C)
    o = _sdpa(qq, kk, vv, heads).reshape(batch, tokens, fq, C).permute(0, 2, 1, 3).reshape(batch * fq * tokens, C)
    out.copy_(o.to(out.dtype))
    return out


def op_im2col_latent(latents, out, *, batch):
    lb, cl, f, h, w = latents.shape
    lat = latents if lb == batch else latents.expand(batch, cl, f, h, w)
    x = lat.permute(0, 2, 1, 3, 4).reshape(batch * f, cl, h, w).float()
    cols = F.unfold(x, 3, padding=1).view(batch * f, cl, 9, h * w)          # [n, c, tap, hw]
    cols = cols.permute(0, 3, 2, 1).reshape(batch * f * h * w, 9 * cl)      # k = tap * cl + c
    out.zero_()
    out[:, :9 * cl].copy_(cols.to(out.dtype))
    return out


def op_timestep_embed(t_table, step, out):
    rows, dim = out.shape
    half = dim // 2
    t = float(t_table[int(step[0])])
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    e = torch.cat([torch.cos(t * freq), torch.sin(t * freq)])
    out.copy_(e.to(out.dtype).unsqueeze(0).expand(rows, dim))
    return out


def op_cfg_ddim_step(model_out, latents, coef, step, *, guidance, v_out=None):
    _, cl, f, h, w = latents.shape
    n = f * h * w
    mo = model_out.float()[:, :cl]
    vu, vc = mo[:n], mo[n:2 * n]
    v = (vu + guidance * (vc - vu)).reshape(f, h, w, cl).permute(3, 0, 1, 2).unsqueeze(0)
    sa, sb, pa, pb = [float(z) for z in coef[int(step[0])]]
    x = latents
    x0 = sa * x - sb * v
    eps = sa * v + sb * x
    latents.copy_(pa * x0 + pb * eps)
    return latents


def op_advance_step(step, n_steps):
    step[0] = (int(step[0]) + 1) % n_steps


def op_tokens_to_bcfhw(x, out):
    b, c, f, h, w = out.shape
    out.copy_(x.float()[:, :c].reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3))
    return out


def op_add(a, b, out):
    out.copy_((a.float() + b.float()).to(out.dtype))
    return out
