import torch
dev=torch.device("cuda:0")
x=torch.empty(1<<30, dtype=torch.float32, device=dev)   # 4 GB
y=torch.empty(1<<30, dtype=torch.float32, device=dev)
def t(fn,reps=5):
  fn(); torch.cuda.synchronize()
  a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b)/reps
ms=t(lambda: x.zero_()); print("memset 4 GB: %.3f ms  %.2f TB/s write"%(ms, 4.295/ms))
ms=t(lambda: x.fill_(1.5)); print("fill   4 GB: %.3f ms  %.2f TB/s write"%(ms, 4.295/ms))
ms=t(lambda: y.copy_(x)); print("copy   4 GB: %.3f ms  %.2f TB/s read+write"%(ms, 8.59/ms))
ms=t(lambda: x.sum()); print("sum    4 GB: %.3f ms  %.2f TB/s read"%(ms, 4.295/ms))
h=torch.empty((196608,5056), dtype=torch.float16, device=dev)
ms=t(lambda: h.fill_(1.0)); print("fill fp16 [196608,5056] (1.99 GB): %.3f ms  %.2f TB/s"%(ms, 1.988/ms))
