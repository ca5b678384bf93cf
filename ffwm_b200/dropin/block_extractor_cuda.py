"""Same surface as the reference's `block_extractor_cuda` pybind module
(cuda/block_extractor/block_extractor_cuda.cc:5-33)."""
from .. import ops


def forward(source, flow_field, output, kernel_size):
    ops.block_extractor_forward(source, flow_field, output, kernel_size)
    return 1


def backward(source, flow_field, grad_output, grad_source, grad_flow_field, kernel_size):
    ops.block_extractor_backward(source, flow_field, grad_output, grad_source, grad_flow_field, kernel_size)
    return 1
