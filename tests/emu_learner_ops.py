"""TEST INFRASTRUCTURE: a torch (CPU, fp32) emulation of the yy_lrn_* kernels with the semantics include/yinyang_b200.h
documents, injected into learner.Learner through its `_ops` test seam so that the layer sequence, the parameter layouts
and the hand-derived backward formulas can be checked against torch autograd on a box without a GPU.  Never imported by
the product."""
import torch


class TorchEmuOps:
    @staticmethod
    def _im2col(X, rows, cols, flip):
        """[P, C] -> [P, 9*C], column t*C + c = X[p + d(t)][c] (zero outside the board), t = kh*3 + kw."""
        P, C = X.shape
        B = P // (rows * cols)
        x = X.reshape(B, rows, cols, C)
        o = X.new_zeros(B, rows, cols, 9, C)
        for t in range(9):
            dx, dy = t // 3 - 1, t % 3 - 1
            if flip:
                dx, dy = -dx, -dy
            xs, xe = max(0, -dx), min(rows, rows - dx)
            ys, ye = max(0, -dy), min(cols, cols - dy)
            o[:, xs:xe, ys:ye, t, :] = x[:, xs + dx:xe + dx, ys + dy:ye + dy, :]
        return o.reshape(P, 9 * C)

    def gemm(self, A, B, C, bias=None, relu=False, accumulate=False, conv=None, bn_sums=None, conv_t=None, bn_bwd=None, b_packed=None):
        Ae = A if conv is None else self._im2col(A, conv[0], conv[1], conv[3])
        Be = B if conv_t is None else self._im2col(B.t().contiguous(), conv_t[0], conv_t[1], False).t()
        r = Ae.double() @ Be.double().t()
        if bias is not None:
            r = r + bias.double()
        if accumulate:
            r = r + C.double()
        C.copy_((r.clamp_min(0) if relu else r).float())
        if bn_sums is not None:
            N = C.shape[1]
            if bn_bwd is None:
                bn_sums[:N] += C.double().sum(0); bn_sums[N:2 * N] += (C.double() ** 2).sum(0)
            else:
                Out, Y, mi = bn_bwd
                dz = (C * (Out > 0)).double()
                xhat = ((Y - mi[:N]) * mi[N:]).double()
                bn_sums[:N] += (dz * xhat).sum(0); bn_sums[N:2 * N] += dz.sum(0)

    def transpose(self, inp, out):
        out.copy_(inp.transpose(-1, -2))

    def im2col_t(self, X, colT, rows, cols):
        colT.copy_(self._im2col(X, rows, cols, False).t())

    def conv_weight_t(self, params, offsets, Wt, cout, cin):
        for l, off in enumerate(offsets.tolist()):
            W = params[off:off + cout * 9 * cin]
            Wt[l].view(cin, 9, cout).copy_(W.view(cout, 9, cin).permute(2, 1, 0))

    def planes_nhwc(self, planes, X0):
        B = planes.shape[0]
        X0.zero_()
        X0.view(B, -1, 8)[:, :, :5] = planes.reshape(B, 5, -1).permute(0, 2, 1)

    def colsum(self, X, out):
        out.copy_(X.double().sum(0).float())

    def bn_forward(self, Y, gamma, beta, residual, out, relu, eps, momentum, ws, mean_invstd, running_mean, running_var, have_sums=False):
        P, C = Y.shape
        if have_sums:
            mean = ws[:C] / P
            var = ws[C:2 * C] / P - mean ** 2
        else:
            mean = Y.double().mean(0)
            var = (Y.double() ** 2).mean(0) - mean ** 2
        mean_invstd[:C] = mean.float(); mean_invstd[C:] = (1.0 / torch.sqrt(var + eps)).float()
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean.float())
            running_var.mul_(1 - momentum).add_(momentum * (var * P / (P - 1)).float())
        o = (Y - mean_invstd[:C]) * mean_invstd[C:] * gamma + beta
        if residual is not None:
            o = o + residual
        out.copy_(o.clamp_min(0) if relu else o)

    def bn_backward(self, dOut, Out, Y, mean_invstd, gamma, ws, dY, dRes, dgamma, dbeta, dbias=None, dYT=None, have_sums=False):
        P, C = Y.shape
        dz = dOut * (Out > 0) if Out is not None else dOut.clone()
        xhat = (Y - mean_invstd[:C]) * mean_invstd[C:]
        if have_sums:                                   # taken by the GEMM that produced dOut
            dg, db = ws[:C].clone(), ws[C:2 * C].clone()
        else:
            dg = (dz.double() * xhat.double()).sum(0); db = dz.double().sum(0)
        dgamma.copy_(dg.float()); dbeta.copy_(db.float())
        dY.copy_(gamma * mean_invstd[C:] * (dz - db.float() / P - xhat * dg.float() / P))
        if dRes is not None:
            dRes.copy_(dz)
        if dbias is not None:
            dbias.add_(dY.double().sum(0).float())
        if dYT is not None:
            dYT.copy_(dY.t())

    def heads_loss(self, logits, pi, h, w2, b2, z, dlogits, dh, dpre, v, dw2, db2, losses):
        B = logits.shape[0]
        lsm = torch.log_softmax(logits, dim=1)
        losses[0] = -(pi * lsm).sum(1).mean()
        dlogits.copy_((lsm.exp() * pi.sum(1, keepdim=True) - pi) / B)
        val = torch.tanh(h.clamp_min(0) @ w2 + b2)
        losses[1] = ((val - z) ** 2).mean()
        dp = 2 * (val - z) / B * (1 - val * val)
        dpre.copy_(dp); v.copy_(val)
        dh.copy_(dp[:, None] * w2[None, :] * (h > 0))
        dw2.copy_(dp @ h.clamp_min(0)); db2.copy_(dp.sum().reshape(1))

    def adam(self, params, grads, m, v, lr, beta1, beta2, eps, wd, step):
        step[0] += 1
        t = int(step[0].item())
        g = grads + wd * params
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
        params.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)
