"""-m gpu: the native C++ host (examples/train_loop.cpp over include/nerfb200.hpp) runs the reference's TrainStep call
sequence — GetGradient with the AcceleratedGradientCalculator callback, optimizer.step, OutputRetriever — on the GPU."""
import re
import subprocess

import pytest

from tests.test_abi_cpu import _build_cpp_example

pytestmark = pytest.mark.gpu


def test_cpp_train_loop_runs_and_loss_is_finite(tmp_path):
    exe = _build_cpp_example(tmp_path)
    r = subprocess.run([str(exe), "", "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    losses = [float(x) for x in re.findall(r"Loss: ([0-9.eE+-]+)", r.stdout)]
    assert len(losses) == 5 and all(0 < l < 10 for l in losses), r.stdout
