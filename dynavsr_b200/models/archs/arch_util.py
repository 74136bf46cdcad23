"""Mirror of codes/models/archs/arch_util.py (hot-path parts: :7-52)."""
import torch.nn as nn
import torch.nn.init as init

from ... import ops


def initialize_weights(net_l, scale=1):
    """Kaiming-normal (fan_in) x ``scale``, zero bias -- same initialisation as arch_util.py:7-24."""
    if not isinstance(net_l, list):
        net_l = [net_l]
    for net in net_l:
        for m in net.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                init.kaiming_normal_(m.weight, a=0, mode='fan_in')
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.zero_()


def make_layer(block, n_layers):
    return nn.Sequential(*[block() for _ in range(n_layers)])


class ResidualBlock_noBN(nn.Module):
    """x + conv2(relu(conv1(x))) (arch_util.py:34-52) as two fused-epilogue kernels on NHWC tensors:
    conv1 carries the ReLU, conv2 carries the identity add."""

    def __init__(self, nf=64):
        super(ResidualBlock_noBN, self).__init__()
        self.conv1 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        initialize_weights([self.conv1, self.conv2], 0.1)

    def forward(self, x):
        out = ops.conv(x, self.conv1.weight, self.conv1.bias, act=ops.ACT_RELU)
        return ops.conv(out, self.conv2.weight, self.conv2.bias, res=x)
