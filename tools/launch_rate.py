"""tools/launch_rate.py -- how many kernel nodes per second does the GPU take from replayed CUDA graphs?  (Both the adaptation pool and
the meta lanes level off near 91 000 library launches per second: is that a launch-rate ceiling?)  Tiny kernels (one 256-thread CTA
each), chains of 2 000 nodes per graph, 1-8 graphs replayed concurrently on their own streams."""
import torch

x = [torch.zeros(256, device='cuda') for _ in range(8)]
streams = [torch.cuda.Stream() for _ in range(8)]
graphs = []
N = 2000
for i in range(8):
    s = streams[i]
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            x[i].add_(1.0)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            x[i].add_(1.0)
    graphs.append(g)
torch.cuda.synchronize()
for P in (1, 2, 4, 6, 8):
    for _ in range(2):
        for i in range(P):
            with torch.cuda.stream(streams[i]):
                graphs[i].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(P):
        streams[i].wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(streams[i]):
            for _ in range(3):
                graphs[i].replay()
    for i in range(P):
        torch.cuda.current_stream().wait_stream(streams[i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('%d concurrent graphs of %d dependent tiny kernels: %.0f kernels/s (%.2f us per kernel per chain)' % (P, N, P * 3 * N / ms * 1e3, ms * 1e3 / (3 * N)))
