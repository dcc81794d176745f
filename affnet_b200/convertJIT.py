"""Stand-ins for the reference's TorchScript exports (convertJIT/AffNetJIT.pt, OriNetJIT.pt, made by
convertJIT/convert_OriNet_and_AffNet_to_JIT.ipynb): modules whose forward returns the RAW head outputs, on the CUDA library.

    aff = AffNetJIT(); aff.load_state_dict(torch.load("pretrained/AffNet.pth")["state_dict"]); aff.eval().cuda()
    xy1 = aff(patches)      # [n,3] = (1 + x0, x1, 1 + x2), what torch.jit.load("AffNetJIT.pt")(patches) returns
"""
from .architectures import AffNetFast, OriNetFast


class AffNetJIT(AffNetFast):
    def __init__(self):
        super().__init__(PS=32)

    def forward(self, input):
        return self.forward_raw(input)


class OriNetJIT(OriNetFast):
    def __init__(self):
        super().__init__(PS=32)

    def forward(self, input):
        return self.forward_raw(input)
