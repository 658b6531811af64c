"""TEST INFRASTRUCTURE ONLY -- exports the oracle YOLOv8n module to ONNX so that OpenCV-DNN
(the CPU stand-in for the reference's TensorRT engine) can run it.
The `onnx` python package is absent offline; the TorchScript exporter only needs it for an
optional post-pass, which is bypassed here."""
import torch


class _Wrapper(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        out = self.m(x)                       # (boxes, scores) or (boxes, scores, keypoints)
        return torch.cat((out[0], out[1]), 2)


def export(model, path, batch=1):
    try:
        from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
        onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    except Exception:
        pass
    x = torch.zeros(batch, 3, 640, 640)
    with torch.no_grad():
        torch.onnx.export(_Wrapper(model).eval(), (x,), path, opset_version=12, dynamo=False,
                          input_names=["images"], output_names=["output"])
    return path
