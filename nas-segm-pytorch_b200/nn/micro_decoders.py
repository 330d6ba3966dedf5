"""Genotype -> decoder graph (reference: src/nn/micro_decoders.py), executed on the fused kernel units of
layer_factory.  Constructor signatures, forward contracts, attribute / state_dict names follow the reference:

  MicroDecoder(inp_sizes, num_classes, config, agg_size=64, num_pools=4, ctx_cell=ContextualCell, aux_cell=False,
               repeats=1)  -> forward(list of feats) -> (out, [aux...])        (micro_decoders.py:142-254)
  TemplateDecoder(inp_sizes, num_classes, config, agg_size=64, num_pools=4, repeats=1, stride_power=1)
               -> forward(list of feats) -> out                                 (micro_decoders.py:257-398)
"""
import torch.nn as nn

from .. import functional as Fn
from ..lib import ACT_NONE
from ..rl.genotypes import AGG_OP_NAMES, OP_NAMES, OP_NAMES_WACV
from .layer_factory import AGG_OPS, OPS, conv3x3, conv_bn_relu


def collect_all(feats, collect_indices, relu=False):
    """Concatenate the loose-end maps at the largest height among them (micro_decoders.py:11-25): the running
    result is resized whenever a taller map arrives, i.e. every map ends up bilinearly resized (possibly in several
    hops) to the tallest size.  Hops are kept so the arithmetic matches the reference exactly."""
    out = [feats[collect_indices[0]]]
    size = tuple(out[0].shape[2:])
    for i in collect_indices[1:]:
        c = feats[i]
        if size[0] > c.shape[2]:
            c = Fn.resize(c, size)
        elif c.shape[2] > size[0]:
            size = tuple(c.shape[2:])
            out = [Fn.resize(o, size) for o in out]
        out.append(c)
    return Fn.concat_resize(out, size, relu)


def _head(module_pre, module_clf, cat):
    """pre_clf (1x1 conv-BN-ReLU) -> conv_clf (3x3, bias); `cat` already carries the F.relu (applied while the concat
    slices were written)."""
    conv, bn = module_pre[0], module_pre[1]
    x = Fn.conv_unit(cat, conv.weight, bn, ks=1, act=Fn.ACT_RELU)
    return clf3x3(module_clf, x)


def clf3x3(conv, x):
    """3x3 classifier conv with bias; logits are always produced in fp32."""
    import torch
    return Fn.conv_unit(x, conv.weight, None, ks=3, stride=conv.stride[0], dil=conv.dilation[0], pad=conv.padding[0],
                        act=ACT_NONE, bias=conv.bias, out_dtype=torch.float32)


class AggregateCell(nn.Module):
    """(optional 1x1 conv-BN-ReLU per input) -> bilinear-up the smaller -> sum (micro_decoders.py:28-51)."""

    def __init__(self, size_1, size_2, agg_size, pre_transform=True):
        super().__init__()
        self.pre_transform = pre_transform
        if pre_transform:
            self.branch_1 = conv_bn_relu(size_1, agg_size, 1, 1, 0)
            self.branch_2 = conv_bn_relu(size_2, agg_size, 1, 1, 0)

    def forward(self, x1, x2):
        if self.pre_transform:
            x1, x2 = Fn.lib.branch_streams.run2(lambda: self.branch_1(x1), lambda: self.branch_2(x2), (x2,))
        s1, s2 = tuple(x1.shape[2:]), tuple(x2.shape[2:])
        if s1 < s2:  # tuple comparison, as torch.Size compares
            return Fn.resize_add(x1, x2)
        return Fn.resize_add(x2, x1)


class ContextualCell(nn.Module):
    """Cell DAG: config = [op0, [pos1, pos2, op1, op2], ...] (micro_decoders.py:54-121)."""

    def __init__(self, config, inp, repeats=1):
        super().__init__()
        self._ops = nn.ModuleList()
        self._pos = []
        self._collect_inds = []
        self._pools = ["x"]
        for ind, op in enumerate(config):
            if ind == 0:
                name = OP_NAMES[op]
                self._ops.append(OPS[name](inp, inp, 1, True, repeats))
                self._pos.append(0)
                self._collect_inds.append(1)
                self._pools.append("{}({})".format(name, self._pools[0]))
                continue
            pos1, pos2, op_id1, op_id2 = op
            for pos, op_id in ((pos1, op_id1), (pos2, op_id2)):
                if pos in self._collect_inds:
                    self._collect_inds.remove(pos)
                name = OP_NAMES[op_id]
                self._ops.append(OPS[name](inp, inp, 1, True, repeats))
                self._pos.append(pos)
                self._pools.append("{}({})".format(name, self._pools[pos]))
            self._ops.append(AggregateCell(size_1=None, size_2=None, agg_size=inp, pre_transform=False))
            self._pos.append([ind * 3 - 1, ind * 3])
            self._collect_inds.append(ind * 3 + 1)
            self._pools.append("sum({},{})".format(self._pools[ind * 3 - 1], self._pools[ind * 3]))

    def forward(self, x):
        feats = [x]
        n, k = len(self._ops), 0
        while k < n:
            pos, op = self._pos[k], self._ops[k]
            if isinstance(pos, list):
                feats.append(op(feats[pos[0]], feats[pos[1]]))
                k += 1
            elif k + 2 < n and not isinstance(self._pos[k + 1], list) and isinstance(self._pos[k + 2], list):
                # the two ops of a cell layer are independent of each other (both read earlier features): concurrent streams
                op2, f1, f2 = self._ops[k + 1], feats[pos], feats[self._pos[k + 1]]
                a, b = Fn.lib.branch_streams.run2(lambda: op(f1), lambda: op2(f2), (f2,))
                feats.extend((a, b))
                k += 2
            else:
                feats.append(op(feats[pos]))
                k += 1
        out = feats[self._collect_inds[0]]
        for i in self._collect_inds[1:]:
            out = Fn.resize_add(feats[i], out)
        return out

    def prettify(self):
        return " + ".join(self._pools[i] for i in self._collect_inds)


class MergeCell(nn.Module):
    """agg(cell(x1), cell(x2)) with separate weights per input (micro_decoders.py:124-139)."""

    def __init__(self, ctx_config, conn, inps, agg_size, ctx_cell, repeats=1):
        super().__init__()
        self.index_1, self.index_2 = conn
        self.op_1 = ctx_cell(ctx_config, inps[0], repeats=repeats)
        self.op_2 = ctx_cell(ctx_config, inps[1], repeats=repeats)
        self.agg = AggregateCell(inps[0], inps[1], agg_size)

    def forward(self, x1, x2):
        a, b = Fn.lib.branch_streams.run2(lambda: self.op_1(x1), lambda: self.op_2(x2), (x2,))
        return self.agg(a, b)

    def prettify(self):
        return self.op_1.prettify()


class MicroDecoder(nn.Module):
    """CVPR-2019 decoder: adaptN 1x1 -> MergeCell per connection (+ auxiliary heads) -> collect -> classifier."""

    def __init__(self, inp_sizes, num_classes, config, agg_size=64, num_pools=4, ctx_cell=ContextualCell,
                 aux_cell=False, repeats=1, **kwargs):
        super().__init__()
        self.aux_cell = aux_cell
        self.agg_size = agg_size
        self.pool = ["l{}".format(i + 1) for i in range(num_pools)]
        for out_idx, size in enumerate(inp_sizes):
            setattr(self, "adapt{}".format(out_idx + 1), conv_bn_relu(size, agg_size, 1, 1, 0, affine=True))
            inp_sizes[out_idx] = agg_size  # the reference overwrites the caller's list too (micro_decoders.py:184)
        sizes = list(inp_sizes)
        cell_config, conns = config
        self.conns, self.ctx, self.repeats, self.ctx_cell = conns, cell_config, repeats, ctx_cell
        self.collect_inds = []
        cells, aux_clfs = [], []
        for block_idx, (ind_1, ind_2) in enumerate(conns):
            for ind in (ind_1, ind_2):
                if ind in self.collect_inds:
                    self.collect_inds.remove(ind)
            cells.append(MergeCell(cell_config, (ind_1, ind_2), (sizes[ind_1], sizes[ind_2]), agg_size, ctx_cell,
                                   repeats=repeats))
            head = nn.Sequential()
            if aux_cell:
                head.add_module("aux_cell", ctx_cell(cell_config, agg_size, repeats=repeats))
            head.add_module("aux_clf", conv3x3(agg_size, num_classes, stride=1, bias=True))
            aux_clfs.append(head)
            self.collect_inds.append(block_idx + num_pools)
            sizes.append(agg_size)
            self.pool.append("({} + {})".format(self.pool[ind_1], self.pool[ind_2]))
        self.cells = nn.ModuleList(cells)
        self.aux_clfs = nn.ModuleList(aux_clfs)
        self.pre_clf = conv_bn_relu(agg_size * len(self.collect_inds), agg_size, 1, 1, 0)
        self.conv_clf = conv3x3(agg_size, num_classes, stride=1, bias=True)
        self.info = " + ".join(self.pool[i] for i in self.collect_inds)
        self.num_classes = num_classes

    def prettify(self, n_params):
        header = "#PARAMS\n\n {:3.2f}M".format(n_params / 1e6)
        return header + "\n\n#Contextual:\n" + self.cells[0].prettify() + "\n\n#Connections:\n" + self.info

    def forward(self, x):
        x = [getattr(self, "adapt{}".format(i + 1))(Fn.as_act(f)) for i, f in enumerate(x)]
        aux_outs = []
        run2 = Fn.lib.branch_streams.run2
        pending, prev = None, None  # the auxiliary head of cell k runs next to cell k+1 (the last one next to the main head)
        for cell, head, conn in zip(self.cells, self.aux_clfs, self.conns):
            if pending is None:
                cell_out = cell(x[conn[0]], x[conn[1]])
            else:
                cell_out, aux = run2(lambda: cell(x[conn[0]], x[conn[1]]), pending, (prev,))
                aux_outs.append(aux)
            x.append(cell_out)
            pending = (lambda h=head, c=cell_out: clf3x3(h.aux_clf, h.aux_cell(c) if self.aux_cell else c))
            prev = cell_out
        main = (lambda: _head(self.pre_clf, self.conv_clf, collect_all(x, self.collect_inds, relu=True)))
        if pending is None:
            return main(), aux_outs
        out, aux = run2(main, pending, (prev,))
        aux_outs.append(aux)
        return out, aux_outs


class TemplateDecoder(nn.Module):
    """WACV-2020 decoder: per structure row (pos1, pos2, cell_id, num_repeats, stride) a template
    [op1(feat1), op2(feat2), agg] is applied num_repeats+1 times with feat1 <- feat2, feat2 <- out in between."""

    def __init__(self, inp_sizes, num_classes, config, agg_size=64, num_pools=4, repeats=1, stride_power=1, **kwargs):
        super().__init__()
        inp_sizes = list(inp_sizes)
        n_scales = len(inp_sizes)
        cells, structure = config
        chans = inp_sizes + [0] * len(structure)
        self._ops = nn.ModuleList()
        self._pos, self._collect_inds, self._repeats = [], [], []
        self._pools = ["l{}".format(j + 1) for j in range(n_scales)]
        for block_idx, (pos1, pos2, cell_id, num_repeats, stride) in enumerate(structure):
            larger = block_idx >= (len(structure) // 2)  # only the first half of the blocks down-samples
            num_repeats += 1
            stride = 2 ** stride
            op_id1, op_id2, op_agg = cells[cell_id]
            ops, pos_list = nn.ModuleList(), []
            new_c, prev_c, agg_c = [0, 0], [0, 0], None
            for rep in range(num_repeats):
                for li, (pos, op_id) in enumerate(((pos1, op_id1), (pos2, op_id2))):
                    if rep == 0:
                        cur = chans[pos]
                        new = cur * int(stride ** stride_power)
                    elif li == 0:
                        cur = new = prev_c[-1]
                    else:
                        cur = new = agg_c
                    new_c[li], prev_c[li] = new, cur
                    if pos in self._collect_inds:
                        self._collect_inds.remove(pos)
                    name = OP_NAMES_WACV[op_id]
                    ops.append(OPS[name](cur, new, stride, True, repeats=repeats))
                    pos_list.append(pos)
                    self._pools.append("{}({})".format(name, self._pools[pos]))
                agg_name = AGG_OP_NAMES[op_agg]
                agg_c = max(new_c)
                ops.append(AGG_OPS[agg_name](new_c[0], new_c[1], agg_c, True, repeats=repeats, larger=larger))
            chans[n_scales + block_idx] = agg_c
            self._pos.append(pos_list)
            self._ops.append(ops)
            self._repeats.append(num_repeats)
            self._collect_inds.append(n_scales + block_idx)
            self._pools.append("{}({},{})".format(agg_name, self._pools[n_scales + block_idx - 2],
                                                  self._pools[n_scales + block_idx - 1]))
        c_pre_clf = sum(c for idx, c in enumerate(chans) if idx in self._collect_inds)
        self.pre_clf = conv_bn_relu(c_pre_clf, agg_size, 1, 1, 0)
        self.conv_clf = conv3x3(agg_size, num_classes, stride=1, bias=True)
        self.info = " + ".join(self._pools[i] for i in self._collect_inds)
        self.num_classes = num_classes
        self.agg_size = agg_size

    def _reset_clf(self, num_classes):
        """Replace the classifier for another label set (micro_decoders.py:367-373; the reference reads self.agg_size there
        without ever setting it -- here it is set, so the call the inference notebooks make works)."""
        if num_classes != self.num_classes:
            dev = self.conv_clf.weight.device
            del self.conv_clf
            self.conv_clf = conv3x3(self.agg_size, num_classes, stride=1, bias=True).to(dev)
            self.num_classes = num_classes

    def prettify(self, n_params):
        return "#PARAMS\n\n {:3.2f}M".format(n_params / 1e6) + "\n\n#Connections:\n" + self.info

    def forward(self, x):
        feats = [Fn.as_act(f) for f in x]
        for pos, ops, repeat in zip(self._pos, self._ops, self._repeats):
            feat1, feat2 = feats[pos[0]], feats[pos[1]]
            out = None
            for i in range(repeat):
                # the two operand branches of an aggregation are independent: concurrent streams (lib._BranchStreams)
                a, b = Fn.lib.branch_streams.run2(lambda: ops[i * 3](feat1), lambda: ops[i * 3 + 1](feat2), (feat2,))
                out = ops[i * 3 + 2](a, b)
                feat1, feat2 = feat2, out
            feats.append(out)
        return _head(self.pre_clf, self.conv_clf, collect_all(feats, self._collect_inds, relu=True))
