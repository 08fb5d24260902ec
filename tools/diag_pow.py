"""Is CPU torch.pow bit-identical to CUDA torch.pow for the sin_cos wave lengths on this host?  (see
oracle/la_oracle.py DIM_MAT_FN)"""
import torch
print("cpu capability:", torch.backends.cpu.get_cpu_capability(), "threads:", torch.get_num_threads())
for fd in (11, 12, 24, 6, 48):
    rng = torch.arange(fd, dtype=torch.float32)
    c = torch.pow(1.0 * 1000, (1.0 / fd) * rng)
    g = torch.pow(1.0 * 1000, (1.0 / fd) * rng.cuda()).cpu()
    d = (c.view(torch.int32) - g.view(torch.int32)).abs()
    print(f"fd={fd}: max ulp diff {int(d.max())}, positions {d.nonzero().flatten().tolist()}")
