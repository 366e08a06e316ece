"""GPU: the bf16 production mode (tcgen05 / TMA convolutions, transposed convolutions, weight gradients) against the
fp32 oracle on a tensor-core-capable small geometry.  Tolerances (SURVEY.md 8(d), bf16 mode): loss within 2e-2
relative; logits within 3e-2 of the max-norm; every parameter gradient has cosine similarity > 0.95 with the oracle's
(bf16 activations quantise at every layer, so element-wise bounds are not meaningful)."""
import pytest
import torch

from util import cuda_net, oracle_net, rel_err

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("tc", [1, 0])
def test_bf16_network_forward_backward(tc):
    from b200unet import ops, synth
    from b200unet.configs import CONFIGS
    from oracle import cl_losses
    geom = CONFIGS["tiny32"]
    data, targets = synth.make_batch(geom)
    onet = oracle_net(geom)
    weights = cl_losses.ds_loss_weights(geom.num_pool)
    oo = onet(data)
    ol = cl_losses.multiple_output_loss2(oo, targets, weights)
    ol.backward()
    ops.set_option("tensor_cores", tc)
    try:
        cnet = cuda_net(geom, onet.state_dict(), precision="bf16")
        co = cnet(data.cuda())
        cl = cl_losses.multiple_output_loss2(co, [t.cuda() for t in targets], weights)
        cl.backward()
        torch.cuda.synchronize()
    finally:
        ops.set_option("tensor_cores", 1)
    assert abs(float(cl) - float(ol)) < 2e-2 * abs(float(ol)), (float(cl), float(ol))
    for a, b in zip(co, oo):
        assert rel_err(a, b) < 3e-2, rel_err(a, b)
    od = dict(onet.named_parameters())
    report, bad = [], []
    for n, p in cnet.named_parameters():
        if od[n].grad is None:
            assert p.grad is None
            continue
        if "conv.bias" in n and "seg" not in n:
            continue
        c = _cos(p.grad, od[n].grad)
        report.append("%-60s cos %.5f" % (n, c))
        if not c > 0.95:
            bad.append(n)
    assert not bad, "\n".join(report)


def test_bf16_tc_step_is_bit_reproducible():
    from b200unet import synth
    from b200unet.configs import CONFIGS
    from oracle import cl_losses
    geom = CONFIGS["tiny32"]
    data, targets = synth.make_batch(geom)
    onet = oracle_net(geom)
    weights = cl_losses.ds_loss_weights(geom.num_pool)
    cnet = cuda_net(geom, onet.state_dict(), precision="bf16")
    ref = None
    for _ in range(3):
        cnet.zero_grad(set_to_none=True)
        out = cnet(data.cuda())
        cl_losses.multiple_output_loss2(out, [t.cuda() for t in targets], weights).backward()
        cur = torch.cat([p.grad.flatten() for p in cnet.parameters() if p.grad is not None]).clone()
        if ref is None:
            ref = cur
        else:
            assert torch.equal(ref.view(torch.int32), cur.view(torch.int32))
