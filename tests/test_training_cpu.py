"""CPU: the flat parameter / gradient layout of the training step (no kernels).  The data-parallel step reduces the
gradients in two buckets -- [0, tail_off) after the backward pass, [tail_off, n) while it is still running -- so the layout
must put exactly the layers whose gradients are final early (from the first Downsample on) behind `tail_off`."""
import torch

from tests.helpers import unet_cfg


def test_store_layout_two_gradient_buckets():
    import tqdne_b200 as tq
    from tqdne_b200 import unet as U
    from tqdne_b200.training import TAIL_FROM_INPUT_BLOCK, _Store

    model = tq.UNetModel(**unet_cfg("1d"))
    st = _Store(model, "cpu")
    assert 0 < st.tail_off < st.n and st.tail_off % 4 == 0
    # every trainable parameter has exactly one 16 B aligned, non-overlapping range
    spans = sorted((off, off + torch.Size(shape).numel()) for off, shape, _ in st.items.values())
    assert all(a % 4 == 0 for a, _ in spans)
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1)) and spans[-1][1] <= st.n
    trainable = [p for p in model.parameters() if p.requires_grad]
    assert len(st.items) == len(trainable)
    # the embedding projections of all ResBlocks come first and adjacent (one dense layer over the concatenated rows)
    res = [m for m in model.modules() if isinstance(m, U.ResBlock)]
    off = 0
    for r in res:
        assert st.items[id(r.emb_layers[1].weight)][0] == off
        off += r.emb_layers[1].weight.numel()
    assert st.emb_rows == sum(r.emb_layers[1].weight.shape[0] for r in res)
    # late bucket = input blocks from the first Downsample on, middle block, output blocks, head -- minus the embedding
    # projections (their gradient is produced last, by the dense-layer backward at the end of the tape)
    emb_ids = {id(p) for r in res for p in r.emb_layers.parameters()}
    tail_mods = list(model.input_blocks)[TAIL_FROM_INPUT_BLOCK:] + [model.middle_block, model.output_blocks, model.out]
    tail_ids = {id(p) for m in tail_mods for p in m.parameters()} - emb_ids
    for p in trainable:
        lo = st.items[id(p)][0]
        assert (lo >= st.tail_off) == (id(p) in tail_ids), "parameter in the wrong gradient bucket"
    # the late bucket holds most of the parameters: what is left to reduce after the backward pass is small
    assert st.n - st.tail_off > 0.8 * st.n
    # views address the flat buffers in the engine's layout
    w = next(m for m in model.modules() if isinstance(m, torch.nn.Conv1d)).weight
    v = st.view(st.G, w)
    assert v.shape == st.items[id(w)][1] and v.data_ptr() == st.G.data_ptr() + 4 * st.items[id(w)][0]
