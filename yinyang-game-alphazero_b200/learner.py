"""Learner step on the GPU: one optimisation step of the reference trainer as hand-written CUDA.

Replaces the body of ``AlphaZeroTrainer.train``'s inner loop (src/yin_yang/ai/trainer.py:120-137)::

    policy_logits, value_preds = self.nnet(boards)            # nnet.train(): batch-norm batch statistics
    total_loss = CrossEntropyLoss(policy_logits, policies) + MSELoss(value_preds.view(-1), values)
    total_loss.backward(); self.optimizer.step()              # Adam(lr=1e-3, weight_decay=1e-4)

for the network of src/yin_yang/ai/neural_network.py:39-123.  There is no torch forward / autograd here: the layer
sequence below calls the ``yy_lrn_*`` entry points of the C ABI (csrc/yy_learn.cu; TF32 tcgen05 GEMMs, float64
batch-norm statistics) on preallocated buffers and the whole step is captured in ONE CUDA graph, replayed per batch.
Parameters, gradients and Adam moments are flat float32 buffers in the kernels' own layouts; ``state_dict()`` /
``load_state_dict()`` convert from / to the reference's tensors (same keys, same shapes), so checkpoints round-trip.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib

HEAD_CH = 32      # policy_conv / value_conv output channels (neural_network.py:58,63)
VALUE_HID = 256   # value_fc1 width (:65)
STEM_CIN = 8      # the 5 input planes padded to 8 channels (16-byte rows)
OP_K, OP_K_CONV, OP_K_CONVT = 0, 1, 2   # YY_OP_* operand modes of yy_lrn_gemm


# --------------------------------------------------------------------------------------------- device ops
def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "row-major 2-D view expected"
    return t.stride(0)


class CudaOps:
    """The yy_lrn_* entry points on torch CUDA tensors (2-D row-major views; outputs written in place)."""

    def __init__(self, precision="3xtf32"):
        if not torch.cuda.is_available():
            raise _lib.YinYangError("the learner needs a CUDA device: there is no CPU fallback")
        self.L = _lib.lib()
        self.precision = {"tf32": 0, "3xtf32": 1}[precision]     # YY_GEMM_TF32 / YY_GEMM_3XTF32
        self.ws = torch.empty(6 << 20, dtype=torch.float32, device="cuda")   # split-K partial tiles (24 MB)
        self.sm_count = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def gemm(self, A, B, C, bias=None, relu=False, accumulate=False, conv=None, bn_sums=None, conv_t=None, bn_bwd=None, b_packed=None):
        """C[M,N] = [C +] A @ B^T (+bias) (relu), B = [N,K].  conv = (rows, cols, cin, flip): A is the activation tensor
        [positions, cin] and the product runs over its implicit im2col (YY_OP_K_CONV).  bn_sums (float64 [>= 2N], zero): also
        accumulate the column sums / sums of squares of C (the statistics of the batch norm that follows).
        conv_t = (rows, cols, cin, 0): B is the TRANSPOSED activation tensor [cin, positions] and stands for its transposed
        im2col [9*cin, positions] (YY_OP_K_CONVT; the weight gradient of a convolution).
        bn_bwd = (Out, Y, mean_invstd) with bn_sums: C is the gradient at Out = relu(bn(Y)); accumulate that layer's
        backward statistics (sum dZ*xhat, sum dZ) into bn_sums.
        b_packed: B as tiled by pack_b (streamed with bulk copies instead of being loaded and split by the producer warps)."""
        M, N = C.shape
        K = B.shape[1]
        assert (B.shape[0] == N or conv_t is not None) and (conv is not None or A.shape == (M, K))
        tiles_m = (M + 127) // 128
        tile_n = min(128, (N + 15) // 16 * 16)
        ctas = tiles_m * ((N + tile_n - 1) // tile_n)
        split = 1
        if ctas < 100:                 # too few tiles to fill 148 SMs: slice K (>= 256 per slice), partial tiles -> workspace
            split = max(1, min(self.sm_count // ctas, K // 256))
        if split > 1 and split * M * N > self.ws.numel():
            split = max(1, self.ws.numel() // (M * N))
        if conv_t is not None:
            split = max(split, (K + 8191) // 8192)
        stats = None
        if bn_sums is not None:
            st = _lib.GemmStats(bn_sums.data_ptr(), None, 0, None, 0, None)
            if bn_bwd is not None:
                o, y, mi = bn_bwd
                st = _lib.GemmStats(bn_sums.data_ptr(), o.data_ptr(), _ld(o), y.data_ptr(), _ld(y), mi.data_ptr())
            stats = ctypes.byref(st)
        cg = conv if conv is not None else conv_t
        geom = ctypes.byref(_lib.ConvGeom(*cg)) if cg is not None else None
        _lib.check(self.L.yy_lrn_gemm(_p(A), _ld(A), OP_K_CONV if conv is not None else OP_K, _p(B), _ld(B),
                                      OP_K_CONVT if conv_t is not None else OP_K, _p(C), _ld(C), M, N, K,
                                      _p(bias), int(relu), int(accumulate), tile_n, split, _p(self.ws), self.ws.numel(), self.precision,
                                      geom, stats, _p(b_packed), self._stream()))

    def packed_b_bytes(self, N, K):
        """Bytes per layer of the packed weight operand, or -1 when this GEMM configuration has none."""
        return int(self.L.yy_lrn_pack_b_bytes(N, K)) if self.precision == 1 else -1

    def pack_b(self, base, offsets, layer_stride, layers, N, K, out):
        _lib.check(self.L.yy_lrn_pack_b(_p(base), _p(offsets), layer_stride, layers, N, K, _p(out), self._stream()))

    def transpose(self, inp, out):
        """out = inp^T; 3-D tensors [batch, R, C] -> [batch, C, R] are transposed matrix by matrix in one launch."""
        if inp.dim() == 3:
            n, R, C = inp.shape
            assert tuple(out.shape) == (n, C, R) and inp.stride(2) == 1 and out.stride(2) == 1
            _lib.check(self.L.yy_lrn_transpose(_p(inp), inp.stride(1), _p(out), out.stride(1), R, C, n, inp.stride(0), out.stride(0), self._stream()))
            return
        R, C = inp.shape
        assert tuple(out.shape) == (C, R)
        _lib.check(self.L.yy_lrn_transpose(_p(inp), _ld(inp), _p(out), _ld(out), R, C, 1, 0, 0, self._stream()))

    def im2col_t(self, X, colT, rows, cols):
        P, C = X.shape
        assert tuple(colT.shape) == (9 * C, P)
        _lib.check(self.L.yy_lrn_im2col_t(_p(X), _ld(X), _p(colT), _ld(colT), P, rows, cols, C, self._stream()))

    def conv_weight_t(self, params, offsets, Wt, cout, cin):
        """Wt[l] = transposed view of the conv weights at params[offsets[l]:] for every layer l (one launch)."""
        _lib.check(self.L.yy_lrn_conv_weight_t(_p(params), _p(offsets), offsets.numel(), _p(Wt), cout, cin, self._stream()))

    def planes_nhwc(self, planes, X0):
        B = planes.shape[0]
        cells = planes.shape[2] * planes.shape[3]
        _lib.check(self.L.yy_lrn_planes_nhwc(_p(planes), _p(X0), B, cells, self._stream()))

    def colsum(self, X, out):
        R, C = X.shape
        _lib.check(self.L.yy_lrn_colsum(_p(X), _ld(X), R, C, _p(out), self._stream()))

    def bn_forward(self, Y, gamma, beta, residual, out, relu, eps, momentum, ws, mean_invstd, running_mean, running_var, have_sums=False):
        P, C = Y.shape
        _lib.check(self.L.yy_lrn_bn_forward(_p(Y), _ld(Y), P, C, _p(gamma), _p(beta), _p(residual), _ld(residual) if residual is not None else 0,
                                            _p(out), _ld(out), int(relu), eps, momentum, _p(ws), int(have_sums), _p(mean_invstd),
                                            _p(running_mean), _p(running_var), self._stream()))

    def bn_backward(self, dOut, Out, Y, mean_invstd, gamma, ws, dY, dRes, dgamma, dbeta, dbias, dYT=None, have_sums=False):
        P, C = Y.shape
        _lib.check(self.L.yy_lrn_bn_backward(_p(dOut), _ld(dOut), _p(Out), _ld(Out) if Out is not None else 0, _p(Y), _ld(Y), P, C,
                                             _p(mean_invstd), _p(gamma), _p(ws), _p(dY), _ld(dY), _p(dRes),
                                             _ld(dRes) if dRes is not None else 0, _p(dgamma), _p(dbeta), _p(dbias), _p(dYT),
                                             _ld(dYT) if dYT is not None else 0, int(have_sums), self._stream()))

    def heads_loss(self, logits, pi, h, w2, b2, z, dlogits, dh, dpre, v, dw2, db2, losses):
        B, A = logits.shape
        _lib.check(self.L.yy_lrn_heads_loss(_p(logits), _ld(logits), _p(pi), A, _p(h), _ld(h), h.shape[1], _p(w2), _p(b2), _p(z), B,
                                            _p(dlogits), _ld(dlogits), _p(dh), _ld(dh), _p(dpre), _p(v), _p(dw2), _p(db2), _p(losses),
                                            self._stream()))

    def adam(self, params, grads, m, v, lr, beta1, beta2, eps, wd, step):
        _lib.check(self.L.yy_lrn_adam(_p(params), _p(grads), _p(m), _p(v), params.numel(), lr, beta1, beta2, eps, wd, _p(step), self._stream()))


# --------------------------------------------------------------------------------------------- parameter layouts
def _conv3_to_kernel(w, cin_pad):          # torch [Cout,Cin,3,3] -> [Cout, 9*cin_pad], column t*cin_pad + ci, t = kh*3+kw
    cout, cin = w.shape[0], w.shape[1]
    k = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    if cin_pad != cin:
        k = torch.cat([k, k.new_zeros(cout, 9, cin_pad - cin)], dim=2)
    return k.reshape(cout, 9 * cin_pad)


def _conv3_from_kernel(k, cin):
    cout = k.shape[0]
    return k.reshape(cout, 3, 3, -1)[..., :cin].permute(0, 3, 1, 2).contiguous()


def _fc_to_kernel(w, cells):               # torch [out, 32*cells] (feature c*cells + hw) -> feature hw*32 + c
    return w.reshape(w.shape[0], HEAD_CH, cells).permute(0, 2, 1).reshape(w.shape[0], HEAD_CH * cells)


def _fc_from_kernel(k, cells):
    return k.reshape(k.shape[0], cells, HEAD_CH).permute(0, 2, 1).reshape(k.shape[0], HEAD_CH * cells).contiguous()


class _Layout:
    """Flat-buffer layout of the trainable parameters (kernel layouts) keyed by the reference's state_dict names."""

    def __init__(self, rows, cols, channels, blocks):
        A = rows * cols
        self.entries = []      # (key, kernel_shape, to_kernel, from_kernel)
        C = channels

        def add(key, shape, to_k=None, from_k=None):
            self.entries.append((key, tuple(shape), to_k or (lambda t: t), from_k or (lambda t: t.clone())))

        def conv3(prefix, cout, cin, cin_pad):
            add(prefix + ".weight", (cout, 9 * cin_pad), lambda t, cp=cin_pad: _conv3_to_kernel(t, cp), lambda k, ci=cin: _conv3_from_kernel(k, ci))
            add(prefix + ".bias", (cout,))

        def bn(prefix, c):
            add(prefix + ".weight", (c,))
            add(prefix + ".bias", (c,))

        conv3("conv1", C, 5, STEM_CIN); bn("bn1", C)
        for b in range(blocks):
            conv3(f"res_blocks.{b}.conv1", C, C, C); bn(f"res_blocks.{b}.bn1", C)
            conv3(f"res_blocks.{b}.conv2", C, C, C); bn(f"res_blocks.{b}.bn2", C)
        for head in ("policy", "value"):
            add(f"{head}_conv.weight", (HEAD_CH, C), lambda t: t.reshape(t.shape[0], t.shape[1]), lambda k: k.reshape(k.shape[0], k.shape[1], 1, 1).clone())
            add(f"{head}_conv.bias", (HEAD_CH,))
            bn(f"{head}_bn", HEAD_CH)
        add("policy_fc.weight", (A, HEAD_CH * A), lambda t: _fc_to_kernel(t, A), lambda k: _fc_from_kernel(k, A))
        add("policy_fc.bias", (A,))
        add("value_fc1.weight", (VALUE_HID, HEAD_CH * A), lambda t: _fc_to_kernel(t, A), lambda k: _fc_from_kernel(k, A))
        add("value_fc1.bias", (VALUE_HID,))
        add("value_fc2.weight", (VALUE_HID,), lambda t: t.reshape(-1), lambda k: k.reshape(1, -1).clone())
        add("value_fc2.bias", (1,))
        self.offsets = {}
        off = 0
        for key, shape, _, _ in self.entries:
            self.offsets[key] = off
            off += (math.prod(shape) + 3) // 4 * 4       # every tensor starts on a 16-byte boundary
        self.total = off

    def bn_prefixes(self, blocks):
        return ["bn1"] + [f"res_blocks.{b}.bn{j}" for b in range(blocks) for j in (1, 2)] + ["policy_bn", "value_bn"]


# --------------------------------------------------------------------------------------------- the learner
class Learner:
    """One Adam step per ``step()`` call on a batch of up to ``batch_size`` samples (the DataLoader's last batch of an
    epoch is smaller, trainer.py:96-100; each batch size gets its own pair of CUDA graphs: forward + backward, Adam).
    With an initialised ``torch.distributed`` process group (and ``data_parallel``) every rank steps on its own batch and
    the gradients are averaged with one all-reduce of the flat buffer between the two graphs."""

    def __init__(self, rows, cols, channels=128, blocks=10, batch_size=64, lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999),
                 eps=1e-8, state_dict=None, device=None, use_graph=True, precision="3xtf32", data_parallel=True, _ops=None):
        # _ops: test seam (tests/ inject a torch emulation of the kernels to check the layer sequence against autograd
        # on a CPU box); the product always runs CudaOps and fails without a CUDA device.
        # precision: "3xtf32" (default; fp32-level GEMMs, what the fp32 reference computes) or "tf32" (single pass)
        self.ops = _ops if _ops is not None else CudaOps(precision)
        self.dev = torch.device(device if device is not None else ("cuda" if _ops is None else "cpu"))
        self.rows, self.cols, self.A = rows, cols, rows * cols
        self.C, self.blocks, self.B = channels, blocks, batch_size
        self.P = batch_size * self.A                 # capacity; b / p below are the current batch's
        self.b, self.p = batch_size, batch_size * self.A
        self.lr, self.wd, self.betas, self.eps = float(lr), float(weight_decay), (float(betas[0]), float(betas[1])), float(eps)
        self.bn_eps, self.bn_momentum = 1e-5, 0.1
        if channels % 4 or 128 % channels:
            raise ValueError("channels must divide 128 and be a multiple of 4")
        if self.A % 4:
            raise ValueError("rows*cols must be a multiple of 4 (16-byte rows for the tensor-core GEMMs)")
        self.layout = _Layout(rows, cols, channels, blocks)
        f32 = dict(dtype=torch.float32, device=self.dev)
        n = self.layout.total
        self.params, self.grads = torch.zeros(n, **f32), torch.zeros(n, **f32)
        self.m, self.v = torch.zeros(n, **f32), torch.zeros(n, **f32)
        self.step_count = torch.zeros(4, dtype=torch.int32, device=self.dev)   # [0] = Adam step; [2..3] kernel scratch
        self._bn = self.layout.bn_prefixes(blocks)
        self.running = {}
        for pre in self._bn:
            c = HEAD_CH if pre in ("policy_bn", "value_bn") else channels
            self.running[pre] = (torch.zeros(c, **f32), torch.ones(c, **f32))
        self.batches_tracked = 0
        self._alloc_buffers()
        self._graphs = {}
        self.world = 1
        if data_parallel:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.world = dist.get_world_size()
        self._use_graph = use_graph and _ops is None
        if state_dict is not None:
            self.load_state_dict(state_dict)
        else:
            self.sync_from_rank0()

    # -- views into the flat buffers
    def _view(self, buf, key):
        for k, shape, _, _ in self.layout.entries:
            if k == key:
                off = self.layout.offsets[key]
                return buf[off:off + math.prod(shape)].view(shape)
        raise KeyError(key)

    def w(self, key):
        return self._view(self.params, key)

    def g(self, key):
        return self._view(self.grads, key)

    def _alloc_buffers(self):
        f32 = dict(dtype=torch.float32, device=self.dev)
        P, C, A, B = self.P, self.C, self.A, self.B
        n_conv = 1 + 2 * self.blocks
        z = lambda *s: torch.zeros(*s, **f32)
        self.planes_in, self.pi_in, self.z_in = z(B, 5, self.rows, self.cols), z(B, A), z(B)
        self.X0 = z(P, STEM_CIN)
        self.Y = [z(P, C) for _ in range(n_conv)]            # convolution outputs (batch-norm inputs), kept for backward
        self.act_all = z(n_conv, P, C)                       # layer outputs (one tensor: their transposes are one launch)
        self.act = [self.act_all[i] for i in range(n_conv)]
        self.mi = {pre: z(2 * (HEAD_CH if pre in ("policy_bn", "value_bn") else C)) for pre in self._bn}
        self.Yh = {h: z(P, HEAD_CH) for h in ("policy", "value")}
        self.acth = {h: z(P, HEAD_CH) for h in ("policy", "value")}
        self.logits, self.hid = z(B, A), z(B, VALUE_HID)     # hid: value_fc1 output BEFORE its ReLU
        self.dlogits, self.dhid, self.dpre, self.v_out = z(B, A), z(B, VALUE_HID), z(B), z(B)
        self.losses = z(2)
        self.bn_ws = torch.zeros(2 * len(self._bn), 2 * 128, dtype=torch.float64, device=self.dev)   # one slot per layer and pass
        self._bn_slot = {pre: i for i, pre in enumerate(self._bn)}
        self.G = [z(P, C) for _ in range(3)]                 # gradients w.r.t. layer outputs (rotating)
        self.dY = z(P, C)
        self.dacth, self.dYh = z(P, HEAD_CH), z(P, HEAD_CH)
        # transposed copies ([channels][positions]) for the weight-gradient GEMMs, whose reduction index is the position
        self.dYT, self.XT, self.XT_all = z(max(C, HEAD_CH) * P), z(C * P), z(n_conv, C, P)
        self.Wt = z(C * HEAD_CH)
        self.Wt_all = z(max(1, 2 * self.blocks), C, 9 * C)           # backward-data views of the tower's conv weights
        pb = self.ops.packed_b_bytes(C, 9 * C) if hasattr(self.ops, "packed_b_bytes") else -1
        self.Bp_fwd = self.Bp_bwd = None                             # tower weights / their backward-data views, tiled for bulk copies
        if pb > 0 and self.blocks:
            self.Bp_fwd = torch.empty(2 * self.blocks, pb, dtype=torch.uint8, device=self.dev)
            self.Bp_bwd = torch.empty(2 * self.blocks, pb, dtype=torch.uint8, device=self.dev)
        self.w_offsets = torch.tensor([self.layout.offsets[f"res_blocks.{k}.conv{j}.weight"] for k in range(self.blocks) for j in (1, 2)],
                                      dtype=torch.int64, device=self.dev)
        kp = (B + 3) // 4 * 4
        self.fcT, self.smallT, self.featT = z(HEAD_CH * A * max(VALUE_HID, A)), z(max(VALUE_HID, A) * kp), z(HEAD_CH * A * kp)

    # -- checkpoints (neural_network.py:198-237 state_dict keys)
    def load_state_dict(self, sd):
        for key, shape, to_k, _ in self.layout.entries:
            t = sd[key].detach().to(self.dev, torch.float32)
            self._view(self.params, key).copy_(to_k(t).reshape(shape))
        for pre in self._bn:
            self.running[pre][0].copy_(sd[pre + ".running_mean"]); self.running[pre][1].copy_(sd[pre + ".running_var"])
        self.batches_tracked = int(sd.get("bn1.num_batches_tracked", 0))
        self.sync_from_rank0()

    def sync_from_rank0(self):
        """Data parallel: every rank adopts rank 0's parameters, Adam moments and batch-norm running statistics (one
        broadcast each), so that the all-reduced gradients are taken at ONE parameter point from the first step on --
        whatever each rank's own initialisation or checkpoint was."""
        if self.world <= 1:
            return
        import torch.distributed as dist
        for buf in (self.params, self.m, self.v):
            dist.broadcast(buf, 0)
        stats = torch.cat([t for pre in self._bn for t in self.running[pre]])
        dist.broadcast(stats, 0)
        off = 0
        for pre in self._bn:
            for t in self.running[pre]:
                t.copy_(stats[off:off + t.numel()]); off += t.numel()
        meta = torch.tensor([self.batches_tracked, int(self.step_count[0].item())], dtype=torch.int64, device=self.dev)
        dist.broadcast(meta, 0)
        self.batches_tracked = int(meta[0].item())
        self.step_count[0] = int(meta[1].item())

    def _export(self, buf):
        out = {}
        for key, shape, _, from_k in self.layout.entries:
            out[key] = from_k(self._view(buf, key)).detach().cpu()
        return out

    def state_dict(self):
        sd = self._export(self.params)
        for pre in self._bn:
            sd[pre + ".running_mean"] = self.running[pre][0].detach().cpu().clone()
            sd[pre + ".running_var"] = self.running[pre][1].detach().cpu().clone()
            sd[pre + ".num_batches_tracked"] = torch.tensor(self.batches_tracked, dtype=torch.long)
        return sd

    def grad_dict(self):
        """Gradients of the last step in the reference's tensor shapes (what ``p.grad`` holds after ``backward()``)."""
        return self._export(self.grads)

    # -- one conv + batch norm (+ skip) + ReLU
    def _geom(self, cin, flip=False):
        return (self.rows, self.cols, cin, int(flip))      # yy_conv_geom: cin = channels of the gathered tensor

    def _conv3_forward(self, x, wkey, bkey, bnpre, li, residual=None):
        ops, P = self.ops, self.p
        sums = self.bn_ws[2 * self._bn_slot[bnpre]]
        ops.gemm(x, self.w(wkey), self.Y[li][:P], bias=self.w(bkey), conv=self._geom(x.shape[1]), bn_sums=sums,
                 b_packed=self.Bp_fwd[li - 1] if (self.Bp_fwd is not None and li >= 1) else None)
        ops.bn_forward(self.Y[li][:P], self.w(bnpre + ".weight"), self.w(bnpre + ".bias"), residual, self.act[li][:P], True, self.bn_eps,
                       self.bn_momentum, sums, self.mi[bnpre], self.running[bnpre][0], self.running[bnpre][1], have_sums=True)

    def _conv3_backward(self, dOut, x_in, wkey, bkey, bnpre, li, dRes, dPrev, accumulate):
        """dOut: gradient w.r.t. act[li].  Writes the layer's parameter gradients; dRes (optional) receives the skip share;
        dPrev (None for the stem) receives / accumulates the gradient w.r.t. x_in."""
        ops, P, C = self.ops, self.p, self.C
        cin = x_in.shape[1]
        dY = self.dY[:P]
        dYT = self.dYT[:C * P].view(C, P)
        ops.bn_backward(dOut, self.act[li][:P], self.Y[li][:P], self.mi[bnpre], self.w(bnpre + ".weight"), self.bn_ws[2 * self._bn_slot[bnpre] + 1],
                        dY, dRes, self.g(bnpre + ".weight"), self.g(bnpre + ".bias"), self.g(bkey), dYT=dYT, have_sums=self._bwd_sums_ready.get(li, False))
        # dW[co][t*cin+ci] = sum_p dY[p][co] * x_in[p + d(t)][ci]: K = positions, both operands from transposed copies
        if li == 0:
            XT = self.XT[:cin * P].view(cin, P)
            ops.transpose(x_in, XT)
        else:
            XT = self._xt_all[li - 1]                                        # transposed in one launch with all layer inputs
        ops.gemm(dYT, XT, self.g(wkey), conv_t=self._geom(cin))              # B = implicit transposed im2col of x_in
        if dPrev is not None:
            # dX[p][ci] = sum_{t,co} dY[p - d(t)][co] * W[co][t*cin+ci]: implicit (mirrored) im2col of dY times Wt
            # ... which is the complete gradient at act[li-1]: that layer's batch-norm backward statistics are taken in the same GEMM
            prev = self._bn[li - 1]
            ops.gemm(dY, self.Wt_all[li - 1], dPrev, accumulate=accumulate, conv=self._geom(C, flip=True),
                     bn_sums=self.bn_ws[2 * self._bn_slot[prev] + 1], bn_bwd=(self.act[li - 1][:P], self.Y[li - 1][:P], self.mi[prev]),
                     b_packed=self.Bp_bwd[li - 1] if self.Bp_bwd is not None else None)
            self._bwd_sums_ready[li - 1] = True

    def _head_forward(self, head, trunk):
        ops, P = self.ops, self.p
        pre = f"{head}_bn"
        sums = self.bn_ws[2 * self._bn_slot[pre]]
        ops.gemm(trunk, self.w(f"{head}_conv.weight"), self.Yh[head][:P], bias=self.w(f"{head}_conv.bias"), bn_sums=sums)
        ops.bn_forward(self.Yh[head][:P], self.w(pre + ".weight"), self.w(pre + ".bias"), None, self.acth[head][:P], True, self.bn_eps,
                       self.bn_momentum, sums, self.mi[pre], self.running[pre][0], self.running[pre][1], have_sums=True)
        return self.acth[head][:P].view(self.b, self.A * HEAD_CH)

    def _kpad(self):
        return (self.b + 3) // 4 * 4                         # the batch as a GEMM K dimension: whole 16-byte chunks

    def _t_batch(self, x, buf):
        """x [b,F] -> [F, kpad] (columns past b zero): operand of a linear layer's weight-gradient GEMM (K = batch)."""
        kp = self._kpad()
        T = buf[:x.shape[1] * kp].view(x.shape[1], kp)
        if kp != self.b:
            T[:, self.b:].zero_()
        self.ops.transpose(x, T[:, :self.b])
        return T

    def _fc_backward(self, dout, feat, wkey, bkey, dfeat):
        """dout [b,N]: gradient at an nn.Linear's output; feat [b,F] its input; writes dW, db and dfeat [b,F] = dout @ W."""
        ops = self.ops
        N, F = dout.shape[1], feat.shape[1]
        ops.gemm(self._t_batch(dout, self.smallT), self._t_batch(feat, self.featT), self.g(wkey))     # dW = dout^T @ feat
        ops.colsum(dout, self.g(bkey))
        WT = self.fcT[:F * N].view(F, N)
        ops.transpose(self.w(wkey), WT)
        ops.gemm(dout, WT, dfeat)

    def _head_backward(self, head, dfeat, trunk, dTrunk, accumulate):
        ops, P, C = self.ops, self.p, self.C
        pre = f"{head}_bn"
        dYh = self.dYh[:P]
        dYhT = self.dYT[:HEAD_CH * P].view(HEAD_CH, P)
        ops.bn_backward(dfeat.view(P, HEAD_CH), self.acth[head][:P], self.Yh[head][:P], self.mi[pre], self.w(pre + ".weight"),
                        self.bn_ws[2 * self._bn_slot[pre] + 1], dYh, None, self.g(pre + ".weight"), self.g(pre + ".bias"), self.g(f"{head}_conv.bias"),
                        dYT=dYhT)
        ops.gemm(dYhT, self._xt_all[2 * self.blocks], self.g(f"{head}_conv.weight"))
        WT = self.Wt[:C * HEAD_CH].view(C, HEAD_CH)
        ops.transpose(self.w(f"{head}_conv.weight"), WT)
        ops.gemm(dYh, WT, dTrunk, accumulate=accumulate)

    def _run(self):
        ops, nb, b, P, C = self.ops, self.blocks, self.b, self.p, self.C
        self.grads.zero_()
        self.bn_ws.zero_()
        self._bwd_sums_ready = {}
        if self.Bp_fwd is not None:                              # the tower's weights, split and tiled once per step
            ops.pack_b(self.params, self.w_offsets, 0, 2 * nb, C, 9 * C, self.Bp_fwd)
        # ---- forward (neural_network.py:94-123, train mode)
        X0 = self.X0[:P]
        ops.planes_nhwc(self.planes_in[:b], X0)
        self._conv3_forward(X0, "conv1.weight", "conv1.bias", "bn1", 0)
        for k in range(nb):
            l1, l2 = 1 + 2 * k, 2 + 2 * k
            pre = f"res_blocks.{k}"
            self._conv3_forward(self.act[l1 - 1][:P], pre + ".conv1.weight", pre + ".conv1.bias", pre + ".bn1", l1)
            self._conv3_forward(self.act[l1][:P], pre + ".conv2.weight", pre + ".conv2.bias", pre + ".bn2", l2, residual=self.act[l1 - 1][:P])
        trunk = self.act[2 * nb][:P]
        fp = self._head_forward("policy", trunk)
        fv = self._head_forward("value", trunk)
        logits, hid = self.logits[:b], self.hid[:b]
        ops.gemm(fp, self.w("policy_fc.weight"), logits, bias=self.w("policy_fc.bias"))
        ops.gemm(fv, self.w("value_fc1.weight"), hid, bias=self.w("value_fc1.bias"))       # its ReLU is applied by the head kernels
        # ---- losses + head gradients (trainer.py:131-133)
        dlogits, dhid = self.dlogits[:b], self.dhid[:b]
        ops.heads_loss(logits, self.pi_in[:b], hid, self.w("value_fc2.weight"), self.w("value_fc2.bias"), self.z_in[:b],
                       dlogits, dhid, self.dpre[:b], self.v_out[:b], self.g("value_fc2.weight"), self.g("value_fc2.bias"), self.losses)
        # ---- backward
        if nb:
            ops.conv_weight_t(self.params, self.w_offsets, self.Wt_all, C, C)
            if self.Bp_bwd is not None:
                ops.pack_b(self.Wt_all, None, C * 9 * C, 2 * nb, C, 9 * C, self.Bp_bwd)
        # every layer output transposed ([C][P]) in one launch: inputs of the weight-gradient GEMMs (the last one = the trunk)
        self._xt_all = self.XT_all.view(-1)[:self.XT_all.shape[0] * C * P].view(-1, C, P)
        ops.transpose(self.act_all[:, :P, :], self._xt_all)
        dfeat = self.dacth[:P].view(b, self.A * HEAD_CH)
        G = [g[:P] for g in self.G]
        self._fc_backward(dlogits, fp, "policy_fc.weight", "policy_fc.bias", dfeat)
        self._head_backward("policy", dfeat, trunk, G[0], accumulate=False)
        self._fc_backward(dhid, fv, "value_fc1.weight", "value_fc1.bias", dfeat)
        self._head_backward("value", dfeat, trunk, G[0], accumulate=True)
        gi = 0                                               # G[gi]: gradient w.r.t. the current block's output
        for k in reversed(range(nb)):
            l1, l2 = 1 + 2 * k, 2 + 2 * k
            pre = f"res_blocks.{k}"
            r, h = (gi + 1) % 3, (gi + 2) % 3
            self._conv3_backward(G[gi], self.act[l1][:P], pre + ".conv2.weight", pre + ".conv2.bias", pre + ".bn2", l2, dRes=G[r], dPrev=G[h], accumulate=False)
            self._conv3_backward(G[h], self.act[l1 - 1][:P], pre + ".conv1.weight", pre + ".conv1.bias", pre + ".bn1", l1, dRes=None, dPrev=G[r], accumulate=True)
            gi = r
        self._conv3_backward(G[gi], X0, "conv1.weight", "conv1.bias", "bn1", 0, dRes=None, dPrev=None, accumulate=False)

    def _run_adam(self):  # trainer.py:52-56
        self.ops.adam(self.params, self.grads, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.step_count)

    def _allreduce_grads(self):
        """Data parallel: one all-reduce (NCCL over NVLink on the GPU box) of the flat gradient buffer, mean over
        ranks -- what DistributedDataParallel does bucket by bucket; batch-norm statistics stay per rank, as in DDP."""
        import torch.distributed as dist
        dist.all_reduce(self.grads)
        self.grads.mul_(1.0 / self.world)

    def step(self, planes, policies, values):
        """planes float32[b,5,n,m], policies float32[b,A], values float32[b] (device tensors, 1 <= b <= batch_size).
        Returns the device tensor [policy_loss, value_loss] of this batch (before the update), like trainer.py:131-132."""
        b = int(planes.shape[0])
        if not 1 <= b <= self.B:
            raise ValueError(f"the learner was built for batches of up to {self.B}, got {b}")
        self.b, self.p = b, b * self.A
        self.planes_in[:b].copy_(planes.reshape(b, 5, self.rows, self.cols)); self.pi_in[:b].copy_(policies); self.z_in[:b].copy_(values.reshape(-1))
        graphs = self._graphs.get(b)
        if graphs is not None:
            graphs[0].replay()
            if self.world > 1:
                self._allreduce_grads()
            graphs[1].replay()
        else:
            self._run()                                      # the first step of a batch size runs eagerly ...
            if self.world > 1:
                self._allreduce_grads()
            self._run_adam()
            if self._use_graph:                              # ... and is then captured (capture does not execute)
                torch.cuda.synchronize()
                g_bwd, g_adam = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_bwd):
                    self._run()
                with torch.cuda.graph(g_adam):
                    self._run_adam()
                self._graphs[b] = (g_bwd, g_adam)
        self.batches_tracked += 1
        return self.losses
