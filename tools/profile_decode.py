"""Target for an ncu launch list of ONE graph-captured greedy decoding step (bf16, Whisper-small decoder, random weights)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ.setdefault("AGA_ALLOW_RANDOM_INIT", "1")
import torch
import aga_b200  # noqa: F401
from aga_b200 import espnet_whisper as EW

dtype = torch.float32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else torch.bfloat16
dec = EW.OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1).cuda().eval()
enc_out = torch.randn(1, 1500, 768, device="cuda", dtype=dtype)
ys = torch.tensor([[50258, 50260, 50259, 50359, 50363]], device="cuda")
gd = dec.greedy_decoder(enc_out, max_len=128)
gd.prefill(ys)
gd.decode(2)
torch.cuda.synchronize()
print("MARK")
gd.decode(2)
torch.cuda.synchronize()
