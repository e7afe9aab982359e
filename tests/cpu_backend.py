"""TEST-ONLY CPU stand-in for libb200gen.so: every C-ABI entry point re-implemented with numpy/torch on HOST pointers,
following include/b200gen.h literally.  Installing it lets the `-m "not gpu"` suite drive the real host code
(generativemodels_b200: modules, weight packing, schedulers, inferers, ctypes marshalling) end to end on a machine
without a GPU.  It is never importable from the product package and proves nothing about the CUDA kernels — those are
covered by the -m gpu tests against the oracle."""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn.functional as F

from generativemodels_b200 import _lib, ops
from tests import igemm_emulator as E


def _obj(ref):
    return ref._obj if hasattr(ref, "_obj") else ref


def _np(ptr, count, ctype):
    if not ptr:
        return None
    return np.ctypeslib.as_array(C.cast(int(ptr), C.POINTER(ctype)), shape=(int(count),))


def f32(ptr, count):
    a = _np(ptr, count, C.c_float)
    return None if a is None else torch.from_numpy(a)


def bf16(ptr, count):
    a = _np(ptr, count, C.c_uint16)
    return None if a is None else torch.from_numpy(a.view(np.int16)).view(ops.H16)


def rows16(ptr, n_rows, pitch, cols):
    """[n_rows, cols] view of 16-bit rows `pitch` elements apart that may be a column slice of a wider buffer: only
    (n_rows - 1) * pitch + cols elements are touched."""
    flat = bf16(ptr, (n_rows - 1) * pitch + cols)
    return torch.as_strided(flat, (n_rows, cols), (pitch, 1))


def i64(ptr, count):
    a = _np(ptr, count, C.c_int64)
    return None if a is None else torch.from_numpy(a)


def _act(x, a):
    if a == _lib.ACT_LEAKYRELU:
        return F.leaky_relu(x, 0.01)
    if a == _lib.ACT_GELU:
        return F.gelu(x)
    if a == _lib.ACT_TANH:
        return torch.tanh(x)
    if a == _lib.ACT_SIGMOID:
        return torch.sigmoid(x)
    return F.relu(x) if a == _lib.ACT_RELU else F.silu(x) if a == _lib.ACT_SILU else x


class FakeLib:
    def b200_igemm(self, p, stream):
        E.emulate(_obj(p))
        return 0

    def b200_groupnorm_workspace_bytes(self, N, spatial, C_):
        return 16

    def b200_nchw_to_nhwc(self, x, N, C_, sp, y, pitch, stream):
        src = f32(x, N * C_ * sp).view(N, C_, sp)
        dst = bf16(y, N * sp * pitch).view(N, sp, pitch)
        dst.zero_()
        dst[:, :, :C_] = src.transpose(1, 2).to(ops.H16)
        return 0

    def b200_nhwc_to_nchw(self, x, dt, N, C_, sp, pitch, y, stream):
        src = (bf16 if dt == _lib.DT_H16 else f32)(x, N * sp * pitch).view(N, sp, pitch)
        f32(y, N * C_ * sp).view(N, C_, sp).copy_(src[:, :, :C_].float().transpose(1, 2))
        return 0

    def _gather_src(self, p):
        xs = []
        for i in range(2):
            if p.x_ptr[i]:
                xs.append(bf16(p.x_ptr[i], p.N * p.spatial * p.x_pitch[i]).view(p.N, p.spatial, p.x_pitch[i])
                          [:, :, : p.x_C[i]].float())
        return torch.cat(xs, 2)

    def b200_groupnorm_stats(self, p, stream):
        p = _obj(p)
        x = self._gather_src(p)                       # [N, S, C]
        N, S, Cc = x.shape
        g = x.view(N, S, p.groups, Cc // p.groups)
        mean = g.mean(dim=(1, 3), keepdim=True)
        var = g.var(dim=(1, 3), unbiased=False, keepdim=True)
        rstd = (var + p.eps).rsqrt()
        gamma, beta = f32(p.gamma, Cc), f32(p.beta, Cc)
        a = (rstd.expand(N, 1, p.groups, Cc // p.groups).reshape(N, Cc)) * gamma
        b = beta - (mean.expand(N, 1, p.groups, Cc // p.groups).reshape(N, Cc)) * a
        f32(p.affine, N * Cc * 2).view(N, Cc, 2).copy_(torch.stack([a, b], -1))
        return 0

    def b200_groupnorm_fused(self, sp, ap, stream):
        sp, ap = _obj(sp), _obj(ap)
        Cc = sp.x_C[0] + (sp.x_C[1] if sp.x_ptr[1] else 0)
        assert Cc % sp.groups == 0 and (not sp.x_ptr[1] or sp.x_C[0] % (Cc // sp.groups) == 0)
        table = torch.empty(sp.N * Cc * 2, dtype=torch.float32)          # the fused kernel keeps this in registers
        sp.affine, ap.affine = table.data_ptr(), table.data_ptr()
        try:
            self.b200_groupnorm_stats(C.byref(sp), stream)
            return self.b200_groupnorm_apply(C.byref(ap), stream)
        finally:
            sp.affine, ap.affine = None, None

    def b200_groupnorm_from_partials(self, p, partial, slots, stream):
        return self.b200_groupnorm_from_partials_ex(p, partial, slots, None, stream)

    def b200_groupnorm_from_partials_ex(self, p, partial, slots, group, stream):
        p = _obj(p)
        ptrs = [int(v) if v else 0 for v in partial]
        sl = [int(v) for v in slots]
        gws = [(int(v) or 8) for v in group] if group is not None else [8, 8]
        N, G = p.N, p.groups
        Cs = [p.x_C[0]] + ([p.x_C[1]] if ptrs[1] else [])
        Cc = sum(Cs)
        cpg = Cc // G
        assert all(cpg % gw == 0 and gw in (8, 4) for gw in gws[:len(Cs)]) and Cs[0] % cpg == 0
        sums = []
        for ptr, nslot, Ci, gw in zip(ptrs, sl, Cs, gws):
            part = f32(ptr, N * nslot * (Ci // gw) * 2).view(N, nslot, Ci // gw, 2).double().sum(1)   # [N, Ci/gw, 2]
            sums.append(part.view(N, Ci // cpg, cpg // gw, 2).sum(2))
        tot = torch.cat(sums, 1)                                                                      # [N, G, 2]
        cnt = float(p.spatial) * cpg
        mean = tot[..., 0] / cnt
        var = (tot[..., 1] / cnt - mean * mean).clamp_min(0)
        rstd = (var + p.eps).rsqrt()
        gamma, beta = f32(p.gamma, Cc), f32(p.beta, Cc)
        a = rstd.float().repeat_interleave(cpg, 1) * gamma
        b = beta - mean.float().repeat_interleave(cpg, 1) * a
        f32(p.affine, N * Cc * 2).view(N, Cc, 2).copy_(torch.stack([a, b], -1))
        return 0

    def b200_spade_apply(self, p, gb, gb_pitch, gb_affine, stream):
        p = _obj(p)
        x = self._gather_src(p)                        # [N, S, C]
        N, S, Cc = x.shape
        ab = f32(p.affine, N * Cc * 2).view(N, 1, Cc, 2)
        g = bf16(gb, N * S * gb_pitch).view(N, S, gb_pitch)[:, :, :2 * Cc].float()
        gab = f32(gb_affine, N * 2 * Cc * 2).view(N, 1, 2 * Cc, 2)
        g = g * gab[..., 0] + gab[..., 1]
        y = _act((x * ab[..., 0] + ab[..., 1]) * (1 + g[..., :Cc]) + g[..., Cc:], p.act)
        dst = bf16(p.y_ptr, N * S * p.y_pitch).view(N, S, p.y_pitch)
        dst.zero_()
        dst[:, :, :Cc] = y.to(ops.H16)
        return 0

    def b200_resize_nearest(self, x, N, D, H, W, pitch, y, OD, OH, OW, stream):
        src = bf16(x, N * D * H * W * pitch).view(N, D, H, W, pitch).float().permute(0, 4, 1, 2, 3)
        out = F.interpolate(src, size=(OD, OH, OW), mode="nearest").permute(0, 2, 3, 4, 1)
        bf16(y, N * OD * OH * OW * pitch).view(N, OD, OH, OW, pitch).copy_(out.to(ops.H16))
        return 0

    def b200_groupnorm_apply(self, p, stream):
        p = _obj(p)
        x = self._gather_src(p)
        N, S, Cc = x.shape
        ab = f32(p.affine, N * Cc * 2).view(N, 1, Cc, 2)
        y = _act(x * ab[..., 0] + ab[..., 1], p.act)
        dst = bf16(p.y_ptr, N * S * p.y_pitch).view(N, S, p.y_pitch)
        dst.zero_()
        dst[:, :, :Cc] = y.to(ops.H16)
        return 0

    def b200_layernorm(self, x, M, C_, xp, g, b, eps, y, yp, stream):
        src = bf16(x, M * xp).view(M, xp)[:, :C_].float()
        out = F.layer_norm(src, (C_,), f32(g, C_), f32(b, C_), eps)
        dst = bf16(y, M * yp).view(M, yp)
        dst.zero_()
        dst[:, :C_] = out.to(ops.H16)
        return 0

    def b200_upsample_nearest2x(self, x, N, D, H, W, pitch, dims, y, stream):
        src = bf16(x, N * D * H * W * pitch).view(N, D, H, W, pitch)
        o = src.repeat_interleave(2, 2).repeat_interleave(2, 3)
        if dims == 3:
            o = o.repeat_interleave(2, 1)
        bf16(y, o.numel()).view(o.shape).copy_(o)
        return 0

    def b200_avgpool2(self, x, N, D, H, W, pitch, dims, y, stream):
        src = bf16(x, N * D * H * W * pitch).view(N, D, H, W, pitch).float().permute(0, 4, 1, 2, 3)
        o = F.avg_pool3d(src, (2, 2, 2) if dims == 3 else (1, 2, 2)).permute(0, 2, 3, 4, 1).contiguous()
        bf16(y, o.numel()).view(o.shape).copy_(o.to(ops.H16))
        return 0

    def b200_axpy_h16(self, a, b, alpha, y, n, stream):
        out = (bf16(a, n).float() + alpha * bf16(b, n).float()).to(ops.H16)
        bf16(y, n).copy_(out)
        return 0

    @staticmethod
    def _geom(g):
        return [int(v) for v in (g if not hasattr(g, "_obj") else g._obj)][:16]

    def b200_tap_gather(self, x, C_, xp, geom, out, op, stream):
        N, D, H, W, OD, OH, OW, kd, kh, kw, sd, sh, sw, pd, ph, pw = self._geom(geom)
        src = bf16(x, N * D * H * W * xp).view(N, D, H, W, xp)[..., :C_].float()
        big = F.pad(src, (0, 0, pw, kw + sw * OW, ph, kh + sh * OH, pd, kd + sd * OD))     # generous high-side zeros
        dst = bf16(out, N * OD * OH * OW * op).view(N, OD, OH, OW, op)
        dst.zero_()
        tap = 0
        for a in range(kd):
            for b in range(kh):
                for c in range(kw):
                    win = big[:, a:a + sd * OD:sd, b:b + sh * OH:sh, c:c + sw * OW:sw, :]
                    dst[..., tap * C_:(tap + 1) * C_] = win.to(ops.H16)
                    tap += 1
        return 0

    def b200_tap_sum(self, y, yp, geom, cout, bias, out, op, odt, stream):
        N, D, H, W, OD, OH, OW, kd, kh, kw, sd, sh, sw, pd, ph, pw = self._geom(geom)
        yy = f32(y, N * D * H * W * yp).view(N, D, H, W, yp)
        big = F.pad(yy, (0, 0, pw, kw + OW, ph, kh + OH, pd, kd + OD))
        acc = torch.zeros(N, OD, OH, OW, cout)
        if bias:
            acc += f32(bias, cout)
        tap = 0
        for a in range(kd):
            for b in range(kh):
                for c in range(kw):
                    acc += big[:, a:a + OD, b:b + OH, c:c + OW, tap * cout:(tap + 1) * cout]
                    tap += 1
        if odt == _lib.DT_H16:
            dst = bf16(out, N * OD * OH * OW * op).view(N, OD, OH, OW, op)
            dst.zero_()
            dst[..., :cout] = acc.to(ops.H16)
        else:
            dst = f32(out, N * OD * OH * OW * op).view(N, OD, OH, OW, op)
            dst.zero_()
            dst[..., :cout] = acc
        return 0

    def b200_copy_channels(self, src, C_, sp, dst, dp, off, rows, stream):
        bf16(dst, rows * dp).view(rows, dp)[:, off:off + C_] = bf16(src, rows * sp).view(rows, sp)[:, :C_]
        return 0

    def b200_geglu(self, x, M, H, xp, y, yp, stream):
        src = bf16(x, M * xp).view(M, xp).float()
        out = src[:, :H] * F.gelu(src[:, H:2 * H])
        dst = bf16(y, M * yp).view(M, yp)
        dst[:, :H] = out.to(ops.H16)
        return 0

    def b200_softmax_rows(self, s, M, S, sp, p, pp, stream):
        sc = f32(s, M * sp).view(M, sp)[:, :S]
        dst = bf16(p, M * pp).view(M, pp)
        dst.zero_()
        dst[:, :S] = torch.softmax(sc, -1).to(ops.H16)
        return 0

    def b200_softmax_rows_partials(self, s, M, S, sp, part, n_tiles, p, pp, stream):
        sc = f32(s, M * sp).view(M, sp)[:, :S]
        pt = f32(part, M * n_tiles * 2).view(M, n_tiles, 2)
        mx = pt[:, :, 0].max(1)[0]
        den = (pt[:, :, 1] * torch.exp(pt[:, :, 0] - mx[:, None])).nan_to_num(0.0).sum(1)
        dst = bf16(p, M * pp).view(M, pp)
        dst.zero_()
        dst[:, :S] = (torch.exp(sc - mx[:, None]) / den[:, None]).to(ops.H16)
        return 0

    def b200_sm_count(self):
        return 2

    def b200_attention_flash_workspace_bytes(self, a):
        return 0

    def b200_igemm_split_workspace_bytes(self, a):
        return 0            # splitting the reduction is a scheduling decision of the CUDA library; results are the same

    def b200_attention_flash(self, a, stream):
        a = _obj(a)
        B, T, S, heads, dh = a.B, a.T, a.S, a.heads, a.dh
        Cc = heads * dh
        qq = rows16(a.q, B * T, a.q_pitch, Cc).float().view(B, T, heads, dh).transpose(1, 2)
        kk = rows16(a.k, B * S, a.k_pitch, Cc).float().view(B, S, heads, dh).transpose(1, 2)
        vv = bf16(a.vt, B * Cc * a.vt_pitch).view(B, Cc, a.vt_pitch)[:, :, :S].float().transpose(1, 2)
        vv = vv.reshape(B, S, heads, dh).transpose(1, 2)
        out = (torch.softmax(a.scale * qq @ kk.transpose(-1, -2), -1) @ vv).transpose(1, 2).reshape(B, T, Cc)
        if a.res:
            out = out + bf16(a.res, B * T * a.res_pitch).view(B, T, a.res_pitch)[:, :, :Cc].float()
        bf16(a.out, B * T * a.out_pitch).view(B, T, a.out_pitch)[:, :, :Cc] = out.to(ops.H16)
        return 0

    def b200_attention_small_ex(self, q, k, v, o, B, T, S, heads, dh, qp, kp, vp, op, scale, kv_rows, causal, q_pos0,
                                pos_dev, stream):
        if pos_dev:
            q_pos0 = int(_np(pos_dev, 1, C.c_int32)[0])
            S = q_pos0 + T
        Cc = heads * dh
        qq = rows16(q, B * T, qp, Cc).float().view(B, T, heads, dh).transpose(1, 2)
        kk = rows16(k, B * kv_rows, kp, Cc).float().view(B, kv_rows, Cc)[:, :S].reshape(B, S, heads, dh).transpose(1, 2)
        vv = bf16(v, B * kv_rows * vp).view(B, kv_rows, vp)[:, :S, :Cc].float().view(B, S, heads, dh).transpose(1, 2)
        sc = scale * qq @ kk.transpose(-1, -2)
        if causal:
            allowed = torch.arange(S)[None, :] <= (q_pos0 + torch.arange(T))[:, None]
            sc = sc.masked_fill(~allowed, float("-inf"))
        out = (torch.softmax(sc, -1) @ vv).transpose(1, 2).reshape(B, T, Cc)
        bf16(o, B * T * op).view(B, T, op)[:, :, :Cc] = out.to(ops.H16)
        return 0

    def b200_rows_linear(self, x, xp, M, K, g, b, eps, w, wp, O, bias, act, res, rp, out, op, odt, stream):
        xs = bf16(x, M * xp).view(M, xp)[:, :K].float()
        if g:
            xs = F.layer_norm(xs, (K,), f32(g, K), f32(b, K), eps).to(ops.H16).float()
        W = bf16(w, O * wp).view(O, wp)[:, :K].float()
        y = xs @ W.t()
        if bias:
            y = y + f32(bias, O)
        y = _act(y, act)
        if res:
            y = y + bf16(res, M * rp).view(M, rp)[:, :O].float()
        if odt == _lib.DT_F32:
            f32(out, M * op).view(M, op)[:, :O] = y
        else:
            bf16(out, M * op).view(M, op)[:, :O] = y.to(ops.H16)
        return 0

    def b200_attention_decode(self, q, k, v, o, B, S, heads, dh, qp, kp, vp, op, scale, kv_rows, pos_dev, stream):
        if pos_dev:
            S = int(_np(pos_dev, 1, C.c_int32)[0]) + 1
        return self.b200_attention_small_ex(q, k, v, o, B, 1, S, heads, dh, qp, kp, vp, op, scale, kv_rows, 0, 0, None,
                                            stream)

    def b200_cache_append(self, src, cache, B, T, L, pitch, pos_dev, stream):
        pos = int(_np(pos_dev, 1, C.c_int32)[0])
        bf16(cache, B * L * pitch).view(B, L, pitch)[:, pos:pos + T] = bf16(src, B * T * pitch).view(B, T, pitch)
        return 0

    def b200_advance_i32(self, p, delta, stream):
        _np(p, 1, C.c_int32)[0] += delta
        return 0

    def b200_embed_tokens(self, tokens, M, seq_len, pos0, tok_emb, pos_emb, C_, out, pitch, pos_dev, stream):
        if pos_dev:
            pos0 = int(_np(pos_dev, 1, C.c_int32)[0])
        tk = i64(tokens, M)
        V, Lmax = int(tk.max()) + 1, pos0 + seq_len
        te = f32(tok_emb, V * C_).view(V, C_)
        pe = f32(pos_emb, Lmax * C_).view(Lmax, C_)
        pos = pos0 + torch.arange(M) % seq_len
        dst = bf16(out, M * pitch).view(M, pitch)
        dst.zero_()
        dst[:, :C_] = (te[tk] + pe[pos]).to(ops.H16)
        return 0

    def b200_attention_small(self, q, k, v, o, B, T, S, heads, dh, qp, kp, vp, op, scale, stream):
        Cc = heads * dh
        qq = bf16(q, B * T * qp).view(B, T, qp)[:, :, :Cc].float().view(B, T, heads, dh).transpose(1, 2)
        kk = bf16(k, B * S * kp).view(B, S, kp)[:, :, :Cc].float().view(B, S, heads, dh).transpose(1, 2)
        vv = bf16(v, B * S * vp).view(B, S, vp)[:, :, :Cc].float().view(B, S, heads, dh).transpose(1, 2)
        out = (torch.softmax(scale * qq @ kk.transpose(-1, -2), -1) @ vv).transpose(1, 2).reshape(B, T, Cc)
        bf16(o, B * T * op).view(B, T, op)[:, :, :Cc] = out.to(ops.H16)
        return 0

    def b200_timestep_embedding(self, t, N, dim, max_period, emb, stream):
        half = dim // 2
        tt = f32(t, N)
        exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32)
        args = tt[:, None] * torch.exp(exponent / half)[None]
        e = torch.cat([torch.cos(args), torch.sin(args)], -1)
        if dim % 2:
            e = F.pad(e, (0, 1))
        f32(emb, N * dim).view(N, dim).copy_(e)
        return 0

    def b200_small_linear(self, x, M, K, W, b, O_, act_in, act_out, y, stream):
        out = F.linear(_act(f32(x, M * K).view(M, K), act_in), f32(W, O_ * K).view(O_, K), f32(b, O_))
        f32(y, M * O_).view(M, O_).copy_(_act(out, act_out))
        return 0

    def b200_ddim_step(self, m, s, nz, c, prev, x0o, n, stream):
        c = _obj(c)
        mm, ss = f32(m, n), f32(s, n)
        if c.prediction_type == _lib.PRED_EPSILON:
            x0, eps = (ss - c.sqrt_beta_prod_t * mm) / c.sqrt_alpha_prod_t, mm
        elif c.prediction_type == _lib.PRED_SAMPLE:
            x0 = mm
            eps = (ss - c.sqrt_alpha_prod_t * x0) / c.sqrt_beta_prod_t
        else:
            x0 = c.sqrt_alpha_prod_t * ss - c.sqrt_beta_prod_t * mm
            eps = c.sqrt_alpha_prod_t * mm + c.sqrt_beta_prod_t * ss
        if c.clip:
            x0 = x0.clamp(c.clip_min, c.clip_max)
        p = c.sqrt_alpha_prod_prev * x0 + c.dir_coef * eps
        if nz:
            p = p + c.sigma * f32(nz, n)
        f32(prev, n).copy_(p)
        if x0o:
            f32(x0o, n).copy_(x0)
        return 0

    def b200_ddpm_step(self, m, s, nz, pv, c, prev, x0o, n, stream):
        c = _obj(c)
        mm, ss = f32(m, n), f32(s, n)
        if c.prediction_type == _lib.PRED_EPSILON:
            x0 = (ss - c.sqrt_beta_prod_t * mm) / c.sqrt_alpha_prod_t
        elif c.prediction_type == _lib.PRED_SAMPLE:
            x0 = mm
        else:
            x0 = c.sqrt_alpha_prod_t * ss - c.sqrt_beta_prod_t * mm
        if c.clip:
            x0 = x0.clamp(c.clip_min, c.clip_max)
        p = c.coef_x0 * x0 + c.coef_xt * ss
        if nz:
            sig = c.sigma
            if c.var_mode == 1:
                sig = f32(pv, n).sqrt()
            elif c.var_mode == 2:
                frac = (f32(pv, n) + 1) / 2
                sig = (frac * c.max_log + (1 - frac) * c.min_log).sqrt()
            p = p + sig * f32(nz, n)
        f32(prev, n).copy_(p)
        if x0o:
            f32(x0o, n).copy_(x0)
        return 0

    def b200_ddpm_kl(self, x0, xt, mo, c, kl_out, ssum, N, per, stream):
        from oracle import torch_oracle as O
        c = _obj(c)
        a, s_, m = (f32(p, N * per).view(N, per) for p in (x0, xt, mo))
        if c.prediction_type == _lib.PRED_EPSILON:
            p0 = (s_ - c.sqrt_beta_prod_t * m) / c.sqrt_alpha_prod_t
        elif c.prediction_type == _lib.PRED_SAMPLE:
            p0 = m
        else:
            p0 = c.sqrt_alpha_prod_t * s_ - c.sqrt_beta_prod_t * m
        if c.clip:
            p0 = p0.clamp(-1, 1)
        pred = c.coef_x0 * p0 + c.coef_xt * s_
        if c.is_t0:
            bw = c.bin_width
            kl = -O.decoder_log_likelihood(a, pred, torch.tensor(0.5 * c.log_pred_var), (0, 1), (0, bw))
        else:
            post = c.coef_x0 * a + c.coef_xt * s_
            kl = 0.5 * (-1.0 + c.log_pred_var - c.log_post_var + math.exp(c.log_post_var - c.log_pred_var)
                        + (post - pred) ** 2 * math.exp(-c.log_pred_var))
        if kl_out:
            f32(kl_out, N * per).view(N, per).copy_(kl)
        _np(ssum, N, C.c_double)[:] += kl.double().sum(1).numpy()
        return 0

    def b200_pndm_step(self, hist, s, c, prev, eps_out, n, stream):
        c = _obj(c)
        e = torch.zeros(n)
        for k in range(c.n_hist):
            e = e + c.w[k] * f32(hist[k], n)
        if eps_out:
            f32(eps_out, n).copy_(e)
        if prev:
            ss = f32(s, n)
            if c.prediction_type == _lib.PRED_V:
                e = c.v_alpha * e + c.v_beta * ss
            f32(prev, n).copy_(c.sample_coeff * ss - c.eps_coeff * e)
        return 0

    def b200_add_noise(self, x0, nz, ca, cb, sign_b, N, per, out, stream):
        a, b = f32(ca, N)[:, None], f32(cb, N)[:, None] * sign_b
        f32(out, N * per).view(N, per).copy_(a * f32(x0, N * per).view(N, per) + b * f32(nz, N * per).view(N, per))
        return 0

    def b200_exp_half_clamped(self, x, lo, hi, y, n, stream):
        f32(y, n).copy_(torch.exp(f32(x, n).clamp(lo, hi) / 2))
        return 0

    def b200_fma_f32(self, a, b, c, y, n, stream):
        f32(y, n).copy_(f32(a, n) + f32(b, n) * f32(c, n))
        return 0

    def b200_scale_f32(self, x, mul, div, y, n, stream):
        f32(y, n).copy_(f32(x, n) * mul / div)
        return 0

    def b200_vq_argmin_gather(self, x, M, D, xp, cb, K, idx, q16, qp, q32, ste, sq, hist, stream):
        xx = f32(x, M * xp).view(M, xp)[:, :D]
        cbk = f32(cb, K * D).view(K, D)
        d = (xx ** 2).sum(1, keepdim=True) + (cbk.t() ** 2).sum(0, keepdim=True) - 2 * xx @ cbk.t()
        ii = torch.max(-d, 1)[1]
        i64(idx, M).copy_(ii)
        qv = cbk[ii]
        if q16:
            dst = bf16(q16, M * qp).view(M, qp)
            dst.zero_()
            dst[:, :D] = qv.to(ops.H16)
        if q32:
            f32(q32, M * D).view(M, D).copy_(xx + (qv - xx) if ste else qv)
        if sq:
            _np(sq, 1, C.c_double)[0] += float(((qv - xx).double() ** 2).sum())
        if hist:
            h = _np(hist, K, C.c_int32)
            h += np.bincount(ii.numpy(), minlength=K).astype(np.int32)
        return 0

    def b200_repack_weight(self, src, cout, cin, taps, transposed, mode, blocks, n_blocks, dst, rows_pad, pitch, stream):
        """include/b200gen.h, b200_repack_weight — literal restatement (fp32 sums in tap order, then one rounding)."""
        w = f32(src, cout * cin * taps)
        w = w.view(cin, cout, taps).transpose(0, 1) if transposed else w.view(cout, cin, taps)      # [co][c][tap]
        out = torch.zeros(rows_pad, pitch, dtype=torch.float32)
        if mode == _lib.REPACK_TAP_IN:
            out[:cout, :taps * cin] = w.permute(0, 2, 1).reshape(cout, taps * cin)
        elif mode == _lib.REPACK_TAP_OUT:
            out[:taps * cout, :cin] = w.permute(2, 0, 1).reshape(taps * cout, cin)
        else:
            for i in range(n_blocks):
                b = blocks[i]
                acc = torch.zeros(cout, b.cs)
                for t in range(b.ntaps):
                    acc = acc + w[:, b.cin0:b.cin0 + b.cs, b.tap[t]]
                out[:cout, b.col0:b.col0 + b.cs] = acc
        if ops.H16 == torch.float16:
            out = out.clamp(-65504.0, 65504.0)
        bf16(dst, rows_pad * pitch).view(rows_pad, pitch).copy_(out.to(ops.H16))
        return 0

    def b200_vq_gather(self, idx, M, cb, K, D, q16, qp, stream):
        dst = bf16(q16, M * qp).view(M, qp)
        dst.zero_()
        dst[:, :D] = f32(cb, K * D).view(K, D)[i64(idx, M)].to(ops.H16)
        return 0


def install(monkeypatch):
    """Route the product's C-ABI calls to the CPU stand-in and lift its CUDA-only guards (tests only)."""
    fake = FakeLib()
    monkeypatch.setattr(_lib, "require_device", lambda: fake)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "igemm_raw", E.emulate)
    import generativemodels_b200.networks._holders as H
    import generativemodels_b200.networks.nets.autoencoderkl as A
    import generativemodels_b200.networks.nets.controlnet as CN
    import generativemodels_b200.networks.nets.diffusion_model_unet as U
    import generativemodels_b200.networks.nets.vqvae as V
    import generativemodels_b200.networks.schedulers.scheduler as S
    import generativemodels_b200.networks.schedulers.ddim as S1
    import generativemodels_b200.networks.schedulers.ddpm as S2
    import generativemodels_b200.networks.schedulers.pndm as S3
    for mod in (H, A, CN, U, V):
        monkeypatch.setattr(mod, "require_cuda", lambda x, m: None, raising=False)

    def prep(*tensors):
        return [None if t is None else (t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous())
                for t in tensors]
    for mod in (S, S1, S2, S3):
        monkeypatch.setattr(mod, "_prep", prep, raising=False)
    return fake
