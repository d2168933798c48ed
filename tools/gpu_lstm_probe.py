"""How fast is the stock nn.LSTM on the reference's [T=26 (as batch), N=256 (as sequence), E=300] input under
different stock backends?  (The LSTM is outside the path; this only informs which stock backend the drop-in calls.)"""
import torch, time
dev = "cuda:0"
torch.manual_seed(0)
lstm = torch.nn.LSTM(input_size=300, hidden_size=1024, num_layers=1, batch_first=True).to(dev)
x = torch.randn(26, 256, 300, device=dev, requires_grad=True)


def step():
    lstm.zero_grad(set_to_none=True)
    o, _ = lstm(x)
    o.sum().backward()


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("cudnn LSTM fwd+bwd: %.3f ms" % timeit(step))
with torch.backends.cudnn.flags(enabled=False):
    print("native LSTM (fp32 matmul) fwd+bwd: %.3f ms" % timeit(step))
    torch.backends.cuda.matmul.allow_tf32 = True
    print("native LSTM (tf32 matmul) fwd+bwd: %.3f ms" % timeit(step))
    torch.backends.cuda.matmul.allow_tf32 = False
o1, _ = lstm(x)
with torch.backends.cudnn.flags(enabled=False):
    o2, _ = lstm(x)
print("max |cudnn - native| = %.3e (|o| max %.3e)" % (float((o1 - o2).abs().max()), float(o1.abs().max())))
