"""CPU: host logic of the throughput body (body.py) -- BatchNorm folding and the walk over the HF ResNet structure
(bottleneck and basic layers, projection shortcuts) reproduce the unmodified modules.  The cuDNN fused ops themselves
are CUDA-only (tests/test_gpu_model.py::test_throughput_body_matches_unfused_modules); here they are replaced by
their definition (conv + bias [+ residual] -> ReLU) to check everything around them."""
import pytest
import torch
import torch.nn.functional as F


@pytest.fixture()
def cpu_fused_ops(monkeypatch):
    from enhance_cb_whisper_b200 import body

    monkeypatch.setattr(body._Conv, "relu", lambda s, x: F.relu(F.conv2d(x, s.w, s.b, s.stride, s.padding)))
    monkeypatch.setattr(body._Conv, "add_relu", lambda s, x, z: F.relu(F.conv2d(x, s.w, s.b, s.stride, s.padding) + z))
    return body


def randomise_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=g)
            m.running_mean = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var = 0.5 + torch.rand(m.running_var.shape, generator=g)


@pytest.mark.parametrize("version", ["resnet-50", "resnet-18"])
def test_folded_body_equals_the_modules(cpu_fused_ops, version):
    from enhance_cb_whisper_b200.model import Resnet, run_body

    torch.manual_seed(0)
    net = Resnet(12, 2, version).eval()
    randomise_bn(net, 1)
    fb = cpu_fused_ops.FusedBody(net, torch.float32)
    x = torch.rand(2, 64, 19, 40)
    with torch.no_grad():
        exp = run_body(net, x)
        h = fb.pool(x)
        for layers in fb.stages:
            h = fb._stage(h, layers)
        got = F.linear(h.mean(dim=(2, 3)), fb.lin_w, fb.lin_b)
    assert got.shape == exp.shape == (2, 2)
    assert (got - exp).abs().max().item() <= 2e-5 * max(1.0, exp.abs().max().item())


def test_fold_matches_conv_then_batchnorm():
    from enhance_cb_whisper_b200 import body

    torch.manual_seed(3)
    conv = torch.nn.Conv2d(5, 7, 3, stride=2, padding=1, bias=True)
    bn = torch.nn.BatchNorm2d(7).eval()
    bn.weight.data, bn.bias.data = torch.rand(7) + 0.5, torch.randn(7)
    bn.running_mean, bn.running_var = torch.randn(7), torch.rand(7) + 0.5
    w, b = body._fold(conv, bn, torch.float32)
    x = torch.randn(2, 5, 9, 11)
    with torch.no_grad():
        assert torch.allclose(F.conv2d(x, w, b, 2, 1), bn(conv(x)), atol=1e-5)


def test_fused_body_refuses_cpu_tensors():
    from enhance_cb_whisper_b200 import body
    from enhance_cb_whisper_b200.model import Resnet

    fb = body.FusedBody(Resnet(12, 2, "resnet-18").eval(), torch.float32)
    with pytest.raises(RuntimeError):
        fb(torch.zeros(1, 64, 8, 8))


def test_score_of_an_empty_batch_is_empty():
    """K = 0 keywords or U = 0 utterances: empty results, no kernel launch (also without a GPU)."""
    import enhance_cb_whisper_b200 as kb

    m = kb.KWSModelB200(n_layers=2, embedding_dim=64, learn_features=False)
    kwd, utt = torch.zeros(0, 2, 10, 64), torch.zeros(3, 2, 40, 64)
    sc, det, lg = m.score(kwd, utt, torch.zeros(0, 2, 10), torch.zeros(3, 2, 40))
    assert sc.shape == (0, 3) and det.shape == (0, 3) and lg.shape == (0, 3, 2) and det.dtype == torch.uint8
    sc, det, lg = m.score_host(utt, kwd, torch.zeros(3, 2, 40), torch.zeros(0, 2, 10), device="cpu")
    assert sc.shape == (3, 0) and lg.shape == (3, 0, 2)
