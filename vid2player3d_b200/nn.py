"""Host side of libb200nn.so (C ABI: include/b200nn.h): the low-level policy MLP and the MVAE mixture-of-experts decoder as
chains of hand-written sm_100a GEMM launches (csrc/b200nn.cu).  Mirrors, by name and argument meaning,

  * the actor of the reference's imitation network (embodied_pose/models/im_network_builder.py:191-230: RunningMeanStd
    normalisation, `actor_mlp` 734 -> 1024 -> 1024 -> 512 with ReLU, linear `mu`) as run by ImitatorPlayer.run_one_step
    (vid2player/players/im_player.py:187-202): `PolicyMLP`;
  * `MixedDecoder.forward(z, c)` (vid2player/motion_vae/model.py:186-252): `MixedDecoder`.

Operands are bf16 with fp32 accumulation (the reference runs the MVAE under autocast, motion_vae/base.py:390-406, and the policy
in fp32; tolerances are stated in tests/test_gpu_nn.py).  Every buffer is allocated once; `forward` is launches only, no
allocation and no host synchronisation, so it can be captured into the step's CUDA graph.  There is no fallback: without the
built library or a CUDA device the constructors raise.
"""
import ctypes as C
import os

import torch

from .build import LIB_NN

ACT = {None: 0, "none": 0, "relu": 1, "elu": 2}
_lib = None

SYMBOLS = ["b200nn_abi_version", "b200nn_last_error", "b200nn_linear_create", "b200nn_linear_destroy", "b200nn_linear_run",
           "b200nn_cast_rows", "b200nn_cast_rows3", "b200nn_cast_rows_masked", "b200nn_gate_softmax"]
ABI_VERSION = 1


class LinearDesc(C.Structure):
    _fields_ = [("a", C.c_void_p), ("lda", C.c_int32), ("w", C.c_void_p), ("ldw", C.c_int32), ("bias", C.c_void_p), ("coef", C.c_void_p),
                ("out", C.c_void_p), ("ldo", C.c_int32), ("out_col0", C.c_int32), ("rows", C.c_int32), ("n", C.c_int32),
                ("n_padded", C.c_int32), ("k_padded", C.c_int32), ("num_experts", C.c_int32), ("act", C.c_int32), ("out_bf16", C.c_int32),
                ("out_min", C.c_float), ("out_max", C.c_float)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_NN):
            raise RuntimeError(f"{LIB_NN} not built - run `python -c 'import __graft_entry__ as g; g.build()'` (no fallback)")
        L = C.CDLL(LIB_NN)
        L.b200nn_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(L, name)
        if L.b200nn_abi_version() != ABI_VERSION:
            raise RuntimeError("libb200nn.so ABI version mismatch - rebuild")
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"b200nn error {rc}: {lib().b200nn_last_error().decode()}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _up(x, m):
    return (x + m - 1) // m * m


def padded_rows(rows):
    return _up(rows, 128)


def padded_bf16(rows, cols, device):
    """zero-initialised bf16 operand buffer [rows padded to 128, cols padded to 64]: the pad stays zero for its lifetime"""
    return torch.zeros(padded_rows(rows), _up(cols, 64), device=device, dtype=torch.bfloat16)


class Linear:
    """out[:, col0:col0+N] = act(sum_e coef[:, e] (a W_e^T + b_e)); one launch.  weight [E, N, K] or [N, K] float (torch layout:
    out x in), bias [E, N] or [N]; `a` / `out` are buffers from padded_bf16 (out may also be a float [rows, N] tensor)."""

    def __init__(self, a, weight, bias, out, rows, act=None, coef=None, out_col0=0, out_clamp=None):
        dev = a.device
        if dev.type != "cuda":
            raise RuntimeError("b200nn runs on a CUDA device only (no CPU fallback)")
        w = weight.detach().to(dev, torch.float32)
        b = bias.detach().to(dev, torch.float32)
        if w.dim() == 2:
            w, b = w[None], b[None]
        E, N, K = w.shape
        bn = 128 if E == 1 else 32
        self.n_padded, self.k_padded = _up(N, bn), _up(K, 64)
        assert a.dtype == torch.bfloat16 and a.is_contiguous() and a.shape[1] >= self.k_padded and a.shape[0] >= padded_rows(rows)
        self.w = torch.zeros(E, self.n_padded, self.k_padded, device=dev, dtype=torch.bfloat16)
        self.w[:, :N, :K] = w.to(torch.bfloat16)
        self.b = torch.zeros(E, self.n_padded, device=dev, dtype=torch.float32)
        self.b[:, :N] = b
        self.a, self.out, self.coef = a, out, coef
        out_bf16 = out.dtype == torch.bfloat16
        assert out.is_contiguous() and out.shape[1] >= out_col0 + N and (out_bf16 or out.dtype == torch.float32)
        assert out.shape[0] >= (padded_rows(rows) if out_bf16 else rows)
        if E > 1:
            assert coef is not None and coef.dtype == torch.float32 and coef.is_contiguous() and tuple(coef.shape) == (rows, E)
        d = LinearDesc(a=a.data_ptr(), lda=a.shape[1], w=self.w.data_ptr(), ldw=self.k_padded, bias=self.b.data_ptr(),
                       coef=coef.data_ptr() if coef is not None else None, out=out.data_ptr(), ldo=out.shape[1], out_col0=out_col0,
                       rows=rows, n=N, n_padded=self.n_padded, k_padded=self.k_padded, num_experts=E, act=ACT[act], out_bf16=int(out_bf16),
                       out_min=out_clamp[0] if out_clamp else 0.0, out_max=out_clamp[1] if out_clamp else 0.0)
        self._h = C.c_void_p()
        _check(lib().b200nn_linear_create(C.byref(d), C.c_int32(dev.index or 0), C.byref(self._h)))
        self.flops = 2.0 * rows * N * K * E

    def run(self):
        _check(lib().b200nn_linear_run(self._h, _stream()))

    def __del__(self):
        try:
            if self._h:
                lib().b200nn_linear_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def cast_rows(src, dst, cols, mean=None, rstd=None, lo=-3.0e38, hi=3.0e38):
    """dst[:, :cols] = bf16(clamp((src - mean) * rstd, lo, hi)); src float [rows, >= cols], dst bf16 buffer"""
    assert src.dtype == torch.float32 and src.stride(-1) == 1 and dst.dtype == torch.bfloat16
    _check(lib().b200nn_cast_rows(C.c_void_p(src.data_ptr()), C.c_int32(src.stride(0)), C.c_void_p(dst.data_ptr()), C.c_int32(dst.shape[1]),
                                  C.c_int32(src.shape[0]), C.c_int32(cols), C.c_void_p(mean.data_ptr()) if mean is not None else None,
                                  C.c_void_p(rstd.data_ptr()) if rstd is not None else None, C.c_float(lo), C.c_float(hi), _stream()))


class PolicyMLP:
    """mu = W4 relu(W3 relu(W2 relu(W1 x + b1) + b2) + b3) + b4 with x = clamp((obs - mean) / sqrt(var + eps), -5, 5)
    (im_network_builder.py:191-230 `actor_mlp` + `mu`, RunningMeanStd; im_player.py:187-202).  `layers` = [(weight [out, in],
    bias [out]), ...] in torch nn.Linear layout; the last layer has no activation.  clamp_actions = 1.0 folds
    ImitatorPlayer's `torch.clamp(action, -1, 1)` (im_player.py:198) into the last layer's epilogue."""

    def __init__(self, layers, num_envs, device, obs_mean=None, obs_var=None, eps=1e-5, clamp_obs=5.0, activation="relu", clamp_actions=None):
        self.rows, self.device = num_envs, torch.device(device)
        self.in_dim = layers[0][0].shape[1]
        self.out_dim = layers[-1][0].shape[0]
        self.clamp_obs = float(clamp_obs)
        self.clamp_actions = clamp_actions
        self.mean = obs_mean.to(self.device, torch.float32).contiguous() if obs_mean is not None else None
        self.rstd = (1.0 / torch.sqrt(obs_var.to(self.device, torch.float32) + eps)).contiguous() if obs_var is not None else None
        if (self.mean is None) != (self.rstd is None):
            raise ValueError("obs_mean and obs_var go together")
        self.x = padded_bf16(num_envs, self.in_dim, self.device)
        self.out = torch.zeros(num_envs, self.out_dim, device=self.device)
        self.layers, a = [], self.x
        for i, (w, b) in enumerate(layers):
            last = i == len(layers) - 1
            o = self.out if last else padded_bf16(num_envs, w.shape[0], self.device)
            self.layers.append(Linear(a, w, b, o, num_envs, act=None if last else activation,
                                      out_clamp=(-clamp_actions, clamp_actions) if (last and clamp_actions) else None))
            a = o
        self.flops = sum(l.flops for l in self.layers)
        self.launches_per_forward = 1 + len(self.layers)

    def forward(self, obs):
        """obs float [rows, in_dim] -> mu float [rows, out_dim] (a static buffer, overwritten by the next call)"""
        cast_rows(obs, self.x, self.in_dim, self.mean, self.rstd, -self.clamp_obs, self.clamp_obs)
        for l in self.layers:
            l.run()
        return self.out

    __call__ = forward

    def forward_prepared(self):
        """the layers only: the normalised / clamped bf16 observation row is already in `self.x` (b200env_obs_imitation_rows writes it
        in the launch that computes the observation)"""
        for l in self.layers:
            l.run()
        return self.out

    @classmethod
    def random(cls, num_envs, device, in_dim=734, units=(1024, 1024, 512), out_dim=75, seed=0, out_gain=0.1, **kw):
        """random weights of the reference's shape (torch nn.Linear default init; the `mu` layer scaled by out_gain)"""
        g = torch.Generator().manual_seed(seed)
        dims = [in_dim, *units, out_dim]
        layers = []
        for i in range(len(dims) - 1):
            bound = 1.0 / dims[i] ** 0.5
            w = (torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * bound
            b = (torch.rand(dims[i + 1], generator=g) * 2 - 1) * bound
            if i == len(dims) - 2:
                w, b = w * out_gain, b * out_gain
            layers.append((w, b))
        return cls(layers, num_envs, device, **kw)


class MixedDecoder:
    """MixedDecoder.forward(z, c) (vid2player/motion_vae/model.py:237-252) with the reference's parameter layout:
    `weights[i]` [E, in_i, out_i], `biases[i]` [E, out_i] for the three decoder layers (input of every layer = cat(z, previous
    output)), `gate` = [(weight [out, in], bias [out])] x 3 (nn.Linear layout, ELU between).  The reference blends the expert weight
    matrices per env and runs a batched GEMV; here every layer is one launch: the E expert products of a 128 x 64 output tile are E
    TMEM accumulators, the softmax coefficients blend them in the epilogue."""

    def __init__(self, weights, biases, gate, num_envs, device, latent_size=32):
        self.rows, self.device, self.L = num_envs, torch.device(device), latent_size
        E = weights[0].shape[0]
        self.E = E
        self.cond = weights[0].shape[1] - latent_size
        hid = weights[0].shape[2]
        self.out_dim = weights[2].shape[2]
        dev = self.device
        self.x0 = padded_bf16(num_envs, latent_size + self.cond, dev)     # [z | c]
        self.x1 = padded_bf16(num_envs, latent_size + hid, dev)           # [z | h1]
        self.x2 = padded_bf16(num_envs, latent_size + hid, dev)           # [z | h2]
        self.g1 = padded_bf16(num_envs, gate[0][0].shape[0], dev)
        self.g2 = padded_bf16(num_envs, gate[1][0].shape[0], dev)
        self.coef = torch.zeros(num_envs, E, device=dev)
        self.out = torch.zeros(num_envs, self.out_dim, device=dev)
        self.gate1 = Linear(self.x0, gate[0][0], gate[0][1], self.g1, num_envs, act="elu")
        self.gate2 = Linear(self.g1, gate[1][0], gate[1][1], self.g2, num_envs, act="elu")
        self.gw = gate[2][0].detach().to(dev, torch.float32).contiguous()
        self.gb = gate[2][1].detach().to(dev, torch.float32).contiguous()
        t = lambda w: w.detach().transpose(1, 2)   # noqa: E731   [E, in, out] -> [E, out, in]
        self.l1 = Linear(self.x0, t(weights[0]), biases[0], self.x1, num_envs, act="elu", coef=self.coef, out_col0=latent_size)
        self.l2 = Linear(self.x1, t(weights[1]), biases[1], self.x2, num_envs, act="elu", coef=self.coef, out_col0=latent_size)
        self.l3 = Linear(self.x2, t(weights[2]), biases[2], self.out, num_envs, act=None, coef=self.coef)
        self.flops = sum(l.flops for l in (self.gate1, self.gate2, self.l1, self.l2, self.l3))
        self.launches_per_forward = 1 + 1 + 2 + 1 + 3

    def forward(self, z, c=None, c_clamp=None):
        """z float [rows, latent], c float [rows, >= cond] -> float [rows, out_dim] (static buffer).  c = None: the condition block
        already sits in the first layer's operand buffer (`set_condition`, or the autoregressive `feed_back`)."""
        assert z.dtype == torch.float32 and z.stride(-1) == 1
        _check(lib().b200nn_cast_rows3(C.c_void_p(z.data_ptr()), C.c_int32(z.stride(0)), C.c_void_p(self.x0.data_ptr()), C.c_void_p(self.x1.data_ptr()),
                                       C.c_void_p(self.x2.data_ptr()), C.c_int32(self.x0.shape[1]), C.c_int32(self.rows), C.c_int32(self.L), _stream()))
        if c is not None:
            self.set_condition(c, c_clamp)
        self.gate1.run()
        self.gate2.run()
        _check(lib().b200nn_gate_softmax(C.c_void_p(self.g2.data_ptr()), C.c_int32(self.g2.shape[1]), C.c_int32(self.gw.shape[1]),
                                         C.c_void_p(self.gw.data_ptr()), C.c_void_p(self.gb.data_ptr()), C.c_int32(self.E),
                                         C.c_void_p(self.coef.data_ptr()), C.c_int32(self.rows), _stream()))
        self.l1.run()
        self.l2.run()
        self.l3.run()
        return self.out

    __call__ = forward

    def set_condition(self, c, clamp=None, row_mask=None):
        """condition block of the first layer's operand <- bf16(clamp(c[:, :cond])); row_mask (bool [rows]): those rows only"""
        lo, hi = (-clamp, clamp) if clamp else (-3.0e38, 3.0e38)
        _cast_cols(c, self.x0, self.L, self.cond, lo, hi, row_mask)

    def feed_back(self, clamp=3.0):
        """autoregression of MVAEPlayer (players/mvae_player.py:201-204): the predicted frame becomes the next condition - one cast
        launch from the output buffer straight into the operand buffer (kept inside the normalised range)"""
        self.set_condition(self.out, clamp)

    @classmethod
    def random(cls, num_envs, device, frame_size=288, latent_size=32, hidden_size=256, num_experts=6, out_extra=2, seed=0):
        """random parameters of the reference's shapes (motion_vae/config.py: latent 32, hidden 256, 6 experts; frame 6 + 24*6 +
        23*3 + 23*3 = 288, + 2 phase outputs): uniform expert weights, bias 0.01, nn.Linear default gate"""
        g = torch.Generator().manual_seed(seed)
        inp, inter, out = latent_size + frame_size, latent_size + hidden_size, frame_size + out_extra
        ws, bs = [], []
        for i, o in ((inp, hidden_size), (inter, hidden_size), (inter, out)):
            bound = (6.0 / i) ** 0.5     # uniform with fan_in = in: activations stay O(1) (the reference's kaiming_uniform_ on the
                                         # 3-d tensor uses fan_in = in * out; training overwrites either)
            ws.append((torch.rand(num_experts, i, o, generator=g) * 2 - 1) * bound)
            bs.append(torch.full((num_experts, o), 0.01))
        gate = []
        for i, o in ((inp, 64), (64, 64), (64, num_experts)):
            bound = 1.0 / i ** 0.5
            gate.append(((torch.rand(o, i, generator=g) * 2 - 1) * bound, (torch.rand(o, generator=g) * 2 - 1) * bound))
        return cls(ws, bs, gate, num_envs, device, latent_size)


def _cast_cols(src, dst, col0, cols, lo=-3.0e38, hi=3.0e38, row_mask=None):
    """dst[:rows, col0:col0+cols] = bf16(clamp(src[:, :cols])) through the strided view (col0 elements into each row of dst)"""
    assert src.dtype == torch.float32 and src.stride(-1) == 1
    view_ptr = dst.data_ptr() + col0 * 2
    if row_mask is not None:
        assert row_mask.dtype == torch.bool and row_mask.is_contiguous() and row_mask.shape[0] == src.shape[0]
        _check(lib().b200nn_cast_rows_masked(C.c_void_p(src.data_ptr()), C.c_int32(src.stride(0)), C.c_void_p(view_ptr), C.c_int32(dst.shape[1]),
                                             C.c_int32(src.shape[0]), C.c_int32(cols), C.c_void_p(row_mask.data_ptr()), C.c_float(lo), C.c_float(hi),
                                             _stream()))
        return
    _check(lib().b200nn_cast_rows(C.c_void_p(src.data_ptr()), C.c_int32(src.stride(0)), C.c_void_p(view_ptr), C.c_int32(dst.shape[1]),
                                  C.c_int32(src.shape[0]), C.c_int32(cols), None, None, C.c_float(lo), C.c_float(hi), _stream()))
