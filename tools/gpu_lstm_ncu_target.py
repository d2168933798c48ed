"""Small target for `ncu`: a few forward+backward passes of the persistent LSTM recurrence at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqa_attention_networks_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
lstm = torch.nn.LSTM(input_size=300, hidden_size=1024, num_layers=1, batch_first=True).to(dev)
x = torch.tanh(torch.randn(256, 26, 300, device=dev)).permute(1, 0, 2).requires_grad_(True)
cot = torch.randn(26, 256, 1024, device=dev)
cache = ops.WeightCache()
for _ in range(4):
    lstm.zero_grad(set_to_none=True)
    o = ops.LstmFn.apply(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, cache)
    (o * cot).sum().backward()
torch.cuda.synchronize()
print("ok")
