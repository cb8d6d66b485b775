"""Host-side weight packer: reference checkpoint -> the engine's bf16 weight image.

Input is the reference's checkpoint format (``torch.save({'state_dict', 'board_size', 'action_size'})``,
src/yin_yang/ai/neural_network.py:198-213) or a bare ``state_dict``.  BatchNorm (eval mode: running
statistics, eps = 1e-5) and the conv bias are folded into each convolution
(neural_network.py:94-123: ``relu(bn(conv(x)))``), weights are rounded to bf16 (nearest-even) and laid out in
the exact order the tower kernel streams them (csrc/yy_nn.cu: stage-ordered no-swizzle K-major core
matrices ``[k/8][out_channel][8]``).  Networks narrower than 128 channels are zero-padded.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

TOWER_C = 128
HEAD_C = 64
BN_EPS = 1e-5


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit patterns (uint16), round to nearest even."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)
    return r.astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def _fold(sd, conv: str, bn: str):
    w = _np(sd[conv + ".weight"]).astype(np.float64)
    b = _np(sd[conv + ".bias"]).astype(np.float64)
    g = _np(sd[bn + ".weight"]).astype(np.float64)
    beta = _np(sd[bn + ".bias"]).astype(np.float64)
    mean = _np(sd[bn + ".running_mean"]).astype(np.float64)
    var = _np(sd[bn + ".running_var"]).astype(np.float64)
    scale = g / np.sqrt(var + BN_EPS)
    return (w * scale[:, None, None, None]).astype(np.float32), ((b - mean) * scale + beta).astype(np.float32)


def _stage_blocks(w2d: np.ndarray, slab: int) -> np.ndarray:
    """w2d [N out][K in] float32 -> concatenated stage blocks, each [slab/8][N][8] bf16 bits."""
    n, k = w2d.shape
    assert k % slab == 0
    bits = to_bf16_bits(w2d).reshape(n, k // slab, slab // 8, 8)     # [oc][stage][kc][j]
    return np.ascontiguousarray(bits.transpose(1, 2, 0, 3)).reshape(-1)  # [stage][kc][oc][j]


FC_PANEL_STAGES = 32


def fc_stream_blocks(wp: np.ndarray, wv: np.ndarray, A: int) -> np.ndarray:
    """policy_fc [A][32A] and value_fc1 [256][32A] -> the FC weight stream of the persistent kernel
    (csrc/yy_tower.cuh make_fc_geo / csrc/yy_fused.cu): order head > K panel (32 stages of K = 64) > M tile (128
    output units; the last policy tile is rounded up to 8 rows) > stage; a stage block is [8][rows][8] bf16 bits."""
    kh = 32 * A
    ks = (kh + 63) // 64
    n_panels = (ks + FC_PANEL_STAGES - 1) // FC_PANEL_STAGES
    out = []
    for w in (wp, wv):
        rows_total = w.shape[0]
        tiles = (rows_total + 127) // 128
        wk = np.zeros((tiles * 128, ks * 64), dtype=np.float32)
        wk[:rows_total, :kh] = w
        for p in range(n_panels):
            s0, s1 = p * FC_PANEL_STAGES, min(ks, (p + 1) * FC_PANEL_STAGES)
            for t in range(tiles):
                r = 128 if t < tiles - 1 or rows_total % 128 == 0 else (rows_total - 128 * t + 7) // 8 * 8
                out.append(_stage_blocks(wk[t * 128: t * 128 + r, s0 * 64: s1 * 64], 64))
    return np.concatenate(out)


def infer_arch(sd):
    channels = int(_np(sd["conv1.weight"]).shape[0])
    blocks = 0
    while f"res_blocks.{blocks}.conv1.weight" in sd:
        blocks += 1
    return channels, blocks


def layout(rows: int, cols: int, channels: int, blocks: int):
    out = (ctypes.c_int64 * 13)()
    _lib.check(_lib.lib().yy_nn_weight_layout(rows, cols, channels, blocks, out))
    names = ["conv_stream", "conv_bias", "fc_policy_w", "fc_policy_b", "fc_value1_w", "fc_value1_b",
             "fc_value2_w", "fc_value2_b", "total", "a_pad", "fc_stream", "fc_stream_bytes", "conv_stream_pair"]
    return dict(zip(names, [int(v) for v in out]))


def pack_state_dict(state_dict, rows: int, cols: int) -> np.ndarray:
    """Returns the weight image as uint8[total] (upload to the device, 256-byte aligned)."""
    sd = state_dict["state_dict"] if "state_dict" in state_dict else state_dict
    channels, blocks = infer_arch(sd)
    if channels > TOWER_C:
        raise ValueError(f"tower kernel supports up to {TOWER_C} channels, checkpoint has {channels}")
    A = rows * cols
    lay = layout(rows, cols, channels, blocks)
    img = np.zeros(lay["total"], dtype=np.uint8)

    def pad_oc_ic(w, oc, ic):
        out = np.zeros((oc, ic) + w.shape[2:], dtype=np.float32)
        out[: w.shape[0], : w.shape[1]] = w
        return out

    stream, biases = [], []
    # stem: 5 planes -> 16 input slots; slots 5/6 carry the bf16 residuals of the row/column fill planes and
    # therefore reuse the weights of slots 3/4 (csrc/yy_nn.cu, input encoder).
    w, b = _fold(sd, "conv1", "bn1")
    w16 = np.zeros((TOWER_C, 16, 3, 3), dtype=np.float32)
    w16[:channels, :5] = w
    w16[:channels, 5] = w[:, 3]
    w16[:channels, 6] = w[:, 4]
    for t in range(9):
        stream.append(_stage_blocks(w16[:, :, t // 3, t % 3], 16))
    biases.append(np.pad(b, (0, TOWER_C - channels)))
    for i in range(blocks):
        for conv, bn in (("conv1", "bn1"), ("conv2", "bn2")):
            w, b = _fold(sd, f"res_blocks.{i}.{conv}", f"res_blocks.{i}.{bn}")
            wp = pad_oc_ic(w, TOWER_C, TOWER_C)
            for t in range(9):
                stream.append(_stage_blocks(wp[:, :, t // 3, t % 3], 64))   # two 64-channel slabs per tap
            biases.append(np.pad(b, (0, TOWER_C - channels)))
    wp_, bp_ = _fold(sd, "policy_conv", "policy_bn")
    wv_, bv_ = _fold(sd, "value_conv", "value_bn")
    wh = np.zeros((HEAD_C, TOWER_C), dtype=np.float32)
    wh[:32, :channels] = wp_[:, :, 0, 0]
    wh[32:, :channels] = wv_[:, :, 0, 0]
    stream.append(_stage_blocks(wh, 64))
    head_bias = np.concatenate([bp_, bv_]).astype(np.float32)

    s = np.concatenate(stream).view(np.uint8)
    assert s.size <= lay["conv_bias"] - lay["conv_stream"], (s.size, lay)
    img[lay["conv_stream"]: lay["conv_stream"] + s.size] = s
    # CTA-pair kernel: a stage is one whole tap (or the whole 1x1 head conv), [kc][oc][8] over ALL its K chunks, split
    # into the two CTAs' halves of the output channels: [half][kc][oc/2][8]  (csrc/yy_tower.cuh layer_info, cg = 2)
    halves = []
    for i, blk in enumerate(stream):
        oc = HEAD_C if i == len(stream) - 1 else TOWER_C
        b3 = blk.reshape(-1, oc, 8)                          # slabs of a tap are consecutive K chunks -> [kc_total][oc][8]
        halves.append(np.ascontiguousarray(np.stack([b3[:, : oc // 2], b3[:, oc // 2:]], axis=0)).reshape(-1))
    s2 = np.concatenate(halves).view(np.uint8)
    assert s2.size == s.size
    img[lay["conv_stream_pair"]: lay["conv_stream_pair"] + s2.size] = s2
    cb = np.concatenate(biases + [head_bias]).astype(np.float32).view(np.uint8)
    img[lay["conv_bias"]: lay["conv_bias"] + cb.size] = cb

    def put(off, arr):
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        img[off: off + raw.size] = raw

    wfp = np.zeros((lay["a_pad"], 32 * A), dtype=np.float32)
    wfp[:A] = _np(sd["policy_fc.weight"]).astype(np.float32)
    put(lay["fc_policy_w"], to_bf16_bits(wfp))
    put(lay["fc_policy_b"], np.pad(_np(sd["policy_fc.bias"]).astype(np.float32), (0, lay["a_pad"] - A)))
    put(lay["fc_value1_w"], to_bf16_bits(_np(sd["value_fc1.weight"]).astype(np.float32)))
    put(lay["fc_value1_b"], _np(sd["value_fc1.bias"]).astype(np.float32))
    put(lay["fc_value2_w"], _np(sd["value_fc2.weight"]).astype(np.float32).reshape(-1))
    put(lay["fc_value2_b"], _np(sd["value_fc2.bias"]).astype(np.float32).reshape(-1))
    fcs = fc_stream_blocks(_np(sd["policy_fc.weight"]).astype(np.float32), _np(sd["value_fc1.weight"]).astype(np.float32), A)
    assert fcs.size * 2 == lay["fc_stream_bytes"], (fcs.size * 2, lay)
    put(lay["fc_stream"], fcs)
    return img


def load_checkpoint(path: str):
    """Reads a reference checkpoint file (neural_network.py:209-213) -> dict with 'state_dict'."""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=True)     # tensors and plain containers only
    if "state_dict" not in ck:
        ck = {"state_dict": ck}
    return ck
