"""-m gpu: every C-ABI kernel against its torch oracle (oracle/ops_ref.py) on seeded inputs."""
import pytest

import kernel_cases as kc


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(kc.ALL_CASES))
def test_kernel_case(name, cuda_device):
    res = kc.ALL_CASES[name](cuda_device)
    assert res.ok, str(res)
