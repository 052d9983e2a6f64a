"""Weight re-layout and gradient un-layout as ONE table-driven kernel launch each (seg3d_gather_pack, csrc/train_ops.cu).

The network's parameters stay what the reference's are - fp32 tensors in nn.Conv3d / nn.ConvTranspose3d layout - and the
kernels read their own layouts (plan.py::_Conv).  A training step changes every weight, so every step re-packs all of them
(forward and data-gradient copies) and un-packs all weight gradients.  Each such copy is a strided gather: described here as
an index map over (dim0, dim1, kd, kh, kw) of the parameter, composed of the same steps _Conv.load takes with torch ops
(zero padding, flip, permute, cast, split), and written into a device table once per plan.
"""
import ctypes

import torch

from . import lib


def flat_layout(named_params, align=4):
    """(offsets by name, total elements) of the parameters laid end to end in `named_params` order, each starting on a
    16-byte boundary - the layout shared by the flat gradient buffer the backward pass returns and the optimiser's state."""
    offsets, off = {}, 0
    for name, p in named_params:
        offsets[name] = off
        off += (p.numel() + align - 1) // align * align
    return offsets, off


class _WView(object):
    """5-D strided window [d0, d1, a, b, c] onto a contiguous fp32 parameter: element = base + sum idx*stride, zero where
    idx[d] >= limit[d] (zero padding)."""

    def __init__(self, shape):
        P0, P1, k = int(shape[0]), int(shape[1]), int(shape[2])
        self.k = k
        self.base = 0
        self.extent = [P0, P1, k, k, k]
        self.limit = [P0, P1, k, k, k]
        self.stride = [P1 * k ** 3, k ** 3, k * k, k, 1]

    def pad0(self, n):                      # zero rows appended to dim 0
        if n > self.extent[0]:
            self.extent[0] = n
        return self

    def flip_taps(self):                    # w.flip(2, 3, 4)
        for d in (2, 3, 4):
            self.base += (self.extent[d] - 1) * self.stride[d]
            self.stride[d] = -self.stride[d]
        return self

    def swap01(self):                       # w.permute(1, 0, 2, 3, 4)
        for lst in (self.extent, self.limit, self.stride):
            lst[0], lst[1] = lst[1], lst[0]
        return self

    def crop0(self, n):                     # w[:n]
        self.extent[0] = min(self.extent[0], n)
        return self


class PackTable(object):
    def __init__(self, device):
        self.device, self.entries, self.counts, self.table, self.block_map = device, [], [], None, None
        self._keep = []

    def add(self, src, dst, size, src_stride, dst_stride, limit=None, src_base=0, dst_base=0, kind=lib.PACK_PLAIN):
        assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
        size = [int(v) for v in size] + [1] * (5 - len(size))
        e = lib.PackEntry()
        e.src, e.dst = src.data_ptr(), dst.data_ptr()
        e.src_base, e.dst_base = int(src_base), int(dst_base)
        for d in range(5):
            e.size[d] = size[d]
            e.limit[d] = int(limit[d]) if limit is not None and d < len(limit) else size[d]
            e.src_stride[d] = int(src_stride[d]) if d < len(src_stride) else 0
            e.dst_stride[d] = int(dst_stride[d]) if d < len(dst_stride) else 0
        e.dtype, e.kind = lib.DTYPE_CODE[dst.dtype], kind
        n = 1
        for v in size:
            n *= v
        assert 0 < n < (1 << 31)
        self.counts.append(n)
        self.entries.append(e)
        self._keep += [src, dst]
        self.table = None

    def add_copy(self, src, dst, n=None, limit=None):
        """dst[:n] = src[:limit] zero-padded (1-D)"""
        n = dst.numel() if n is None else n
        self.add(src, dst, [1, 1, 1, 1, n], [0, 0, 0, 0, 1], [0, 0, 0, 0, 1], limit=[1, 1, 1, 1, n if limit is None else limit])

    def add_view(self, src, view, order, dst, dst_shape=None, dst_stride=None, dst_base=0, kind=lib.PACK_PLAIN):
        """dst (row-major over view dims taken in `order`, or with explicit dst strides) = view"""
        size = [view.extent[d] for d in order]
        limit = [view.limit[d] for d in order]
        sstride = [view.stride[d] for d in order]
        if dst_stride is None:
            dst_stride, acc = [0] * 5, 1
            for i in range(4, -1, -1):
                dst_stride[i] = acc
                acc *= size[i]
        self.add(src, dst, size, sstride, dst_stride, limit, view.base, dst_base, kind)

    def run(self):
        if not self.entries:
            return
        if self.table is None:
            arr = (lib.PackEntry * len(self.entries))(*self.entries)
            raw = bytes(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr)))
            self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
            pairs = [(i, c) for i, n in enumerate(self.counts) for c in range((n + lib.PACK_CHUNK - 1) // lib.PACK_CHUNK)]
            self.block_map = torch.tensor(pairs, dtype=torch.int32).reshape(-1).to(self.device)
        lib.call('seg3d_gather_pack', lib.ptr(self.table), lib.ptr(self.block_map), self.block_map.numel() // 2, lib.stream_ptr())


def add_conv_pack(table, conv, weight, bias):
    """entries that rebuild conv.w (and conv.w_fold, conv.bias) from the parameter tensors - the index-map form of
    plan.py::_Conv.load"""
    v = _WView(weight.shape)
    if conv.pad_dim0:
        v.pad0(conv.pad_dim0)
    if conv.transform is not None and getattr(conv.transform, 'seg3d_kind', None) == 'k3_dgrad':
        v.flip_taps().swap01()
    elif conv.transform is not None and getattr(conv.transform, 'seg3d_kind', None) != 'identity':
        raise NotImplementedError('no index map for this weight transform')
    T = conv.mode == lib.CONV_T2S2
    if not T and conv.cout != conv.real_cout:
        v.pad0(conv.cout)                    # zero output channels up to the tensor-core minimum
    # dims of v: non-transposed [Cout, Cin, a, b, c]; transposed conv [Cin, Cout, a, b, c]
    if conv.impl == lib.IMPL_SIMT:
        order = (0, 2, 3, 4, 1) if T else (2, 3, 4, 1, 0)          # [Cin][tap][Cout]  /  [tap][Cin][Cout]
        table.add_view(weight, v, order, conv.w)
    else:
        order = (2, 3, 4, 1, 0) if T else (2, 3, 4, 0, 1)          # [tap][Cout][Cin]
        if conv.split:
            size = [v.extent[d] for d in order]
            K = size[4]
            ds, acc = [0] * 5, 2 * K
            ds[4] = 1
            for i in range(3, -1, -1):
                ds[i] = acc
                acc *= size[i]
            table.add_view(weight, v, order, conv.w, dst_stride=ds, dst_base=0, kind=lib.PACK_SPLIT_HI)
            table.add_view(weight, v, order, conv.w, dst_stride=ds, dst_base=K, kind=lib.PACK_SPLIT_LO)
        else:
            table.add_view(weight, v, order, conv.w)
    if conv.fold:
        C = conv.real_cout
        f = _WView(weight.shape).crop0(C)
        NP, row = conv.w_fold.shape[1], conv.w_fold.shape[2]
        # wf[kd][(kh, kw, co)][ci]; rows 9C..NP stay zero.  Split mode: rows are [whi(Cin) | wlo(Cin)]
        ds = [NP * row, 3 * C * row, C * row, row, 1]
        if conv.split:
            table.add_view(weight, f, (2, 3, 4, 0, 1), conv.w_fold, dst_stride=ds, dst_base=0, kind=lib.PACK_SPLIT_HI)
            table.add_view(weight, f, (2, 3, 4, 0, 1), conv.w_fold, dst_stride=ds, dst_base=row // 2, kind=lib.PACK_SPLIT_LO)
        else:
            table.add_view(weight, f, (2, 3, 4, 0, 1), conv.w_fold, dst_stride=ds)
    if conv.bias is not None and bias is not None:
        table.add_copy(bias, conv.bias, conv.bias.numel(), limit=bias.numel())
