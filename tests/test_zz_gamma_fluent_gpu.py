"""GPU parity of the Gamma_Fluent_Toy fixture (Gamma-family distributions with fluent shapes, sampled
in the step): both CUDA back-ends against the oracle.  Kept in its own file, collected last."""
import pytest

from .parity import check_forward, check_gradient, forward_case, gradient_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backend", ["cuda", "jit"])
def test_gamma_fluent_cuda(cuda_device, backend):
    check_forward(forward_case(backend, "gammafluent", B=37, T=5, relaxed=False))
    c = gradient_case(backend, "gammafluent", B=29, T=4)
    check_forward(c)
    check_gradient(c)
