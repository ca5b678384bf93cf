"""Same surface as the reference's `resample2d_cuda` pybind module
(cuda/resample2d_package/resample2d_cuda.cc:6-33): in-place on caller-owned
tensors, returns 1."""
from .. import ops


def forward(input1, input2, output, kernel_size, dilation):
    ops.resample2d_forward(input1, input2, output, kernel_size, dilation)
    return 1


def backward(input1, input2, gradOutput, gradInput1, gradInput2, kernel_size, dilation):
    ops.resample2d_backward(input1, input2, gradOutput, gradInput1, gradInput2, kernel_size, dilation)
    return 1
