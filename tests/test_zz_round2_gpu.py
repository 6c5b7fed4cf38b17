"""GPU tests written after round 1's GPU budget was spent: they exercise code that is built only from kernels
already validated on a B200 (tests/test_conv_ops.py, tests/test_convnet_trainer.py) but has not itself run on one
yet.  The file sorts last on purpose, so that under ``pytest -x`` a surprise here cannot mask the validated suites.
"""
import pytest
import torch

from colearn_federated_learning_b200.fl.convnet import ConvNetTrainer
from colearn_federated_learning_b200.fl.evaluate import evaluate, predict
from colearn_federated_learning_b200.models.registry import flatten_params
from colearn_federated_learning_b200.models.resnet import ResNet18

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_resnet_eval_mode_inference_on_own_kernels():
    dev = _dev()
    torch.manual_seed(3)
    model = ResNet18(10)
    with torch.no_grad():
        for m in model.modules():
            if hasattr(m, "running_mean"):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    x = torch.randn(128 + 37, 3, 32, 32)
    model.eval()
    with torch.no_grad():
        want = model(x)
    model.to(dev)
    flat = flatten_params(model)
    tr = ConvNetTrainer.cached(model, flat, 128, (32, 32))
    tr.load(flat, model)
    got = tr.infer(x.to(dev)).cpu()
    assert float((got - want).abs().max()) < 0.15 * float(want.abs().max()) + 0.05
    assert float((got.argmax(1) == want.argmax(1)).float().mean()) > 0.9
    # public helpers route conv nets on a GPU through the same path
    pred = predict(model, x.to(dev), flat).cpu().view(-1)
    assert float((pred == want.argmax(1)).float().mean()) > 0.9
    res = evaluate(model, x.to(dev), want.argmax(1).to(dev), flat, verbose=False)
    assert res["accuracy"] > 0.9 and res["n"] == x.shape[0]
