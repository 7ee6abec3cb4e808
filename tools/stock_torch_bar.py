"""The practical bar: FC-DenseNet67 written with STOCK torch.nn modules (nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d /
nn.MaxPool2d / nn.Dropout2d, cuDNN kernels), channels_last, bf16 autocast, cudnn.benchmark, class weights computed on the
device (no host sync), AdamW(fused=True), optionally CUDA-graphed — i.e. what a competent PyTorch user gets on the same
B200 without writing a kernel.  Same architecture and loss as the reference (models/FCDenseNet/tiramisu.py:21-125,
trainingModules/SimpleTrain.py:11-30); random init, synthetic batch.  Prints one JSON line; bench.py's value divided by
these numbers is the honest speed-up (the CPU ratio says nothing about kernel quality).

    python tools/stock_torch_bar.py [train|infer] [--graph]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F

import bench


class DenseLayer(nn.Sequential):
    def __init__(self, cin, g):
        super().__init__(nn.BatchNorm2d(cin), nn.ReLU(True), nn.Conv2d(cin, g, 3, padding=1), nn.Dropout2d(0.2))


class DenseBlock(nn.Module):
    def __init__(self, cin, g, n, only_new):
        super().__init__()
        self.layers = nn.ModuleList([DenseLayer(cin + i * g, g) for i in range(n)])
        self.only_new = only_new

    def forward(self, x):
        new = []
        for l in self.layers:
            o = l(x)
            x = torch.cat([x, o], 1)
            new.append(o)
        return torch.cat(new, 1) if self.only_new else x


class FCDenseNet67(nn.Module):
    def __init__(self, ncls=4, g=16, n=5):
        super().__init__()
        self.first = nn.Conv2d(3, 48, 3, padding=1)
        cur, skips = 48, []
        self.down, self.td = nn.ModuleList(), nn.ModuleList()
        for _ in range(5):
            self.down.append(DenseBlock(cur, g, n, False))
            cur += g * n
            skips.insert(0, cur)
            self.td.append(nn.Sequential(nn.BatchNorm2d(cur), nn.ReLU(True), nn.Conv2d(cur, cur, 1), nn.Dropout2d(0.2), nn.MaxPool2d(2)))
        self.bott = DenseBlock(cur, g, n, True)
        prev = g * n
        self.tu, self.up = nn.ModuleList(), nn.ModuleList()
        for i in range(5):
            self.tu.append(nn.ConvTranspose2d(prev, prev, 3, stride=2))
            self.up.append(DenseBlock(prev + skips[i], g, n, i != 4))
            prev = g * n
        self.final = nn.Conv2d(288, ncls, 1)

    def forward(self, x):
        x = self.first(x)
        skips = []
        for d, t in zip(self.down, self.td):
            x = d(x)
            skips.append(x)
            x = t(x)
        x = self.bott(x)
        for t, u in zip(self.tu, self.up):
            s = skips.pop()
            x = t(x)[:, :, :s.shape[2], :s.shape[3]]
            x = u(torch.cat([x, s], 1))
        x = F.normalize(x.float(), dim=1)
        return torch.softmax(self.final(x) / 0.05, 1)


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "train"
    use_graph = "--graph" in sys.argv
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda")
    torch.manual_seed(0)
    net = FCDenseNet67().to(dev).to(memory_format=torch.channels_last)
    out = {"what": "stock torch.nn FC-DenseNet67, channels_last, bf16 autocast, cudnn.benchmark", "mode": mode, "cuda_graph": use_graph}
    if mode == "train":
        B, H, W = 64, 120, 160
        x, y = bench.synthetic_batch(B, H, W, 4, 1)
        x, y = x.to(dev).contiguous(memory_format=torch.channels_last), y.to(dev)
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=use_graph)
        net.train()

        def step():
            opt.zero_grad(set_to_none=False)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                p = net(x)
            w = torch.reciprocal(F.one_hot(y, 4).sum((0, 1, 2)).float())   # device-side class weights, no host sync
            loss = F.cross_entropy(p.float(), y, weight=w)
            loss.backward()
            opt.step()
            return loss
        if use_graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            fn = g.replay
        else:
            fn = step
        ms = timed(fn, 10)
        out.update(images_per_s=round(B / (ms / 1e3), 1), ms_per_step=round(ms, 3), batch=B, hw=[H, W])
    else:
        B, H, W = 32, 480, 640
        x, y = bench.synthetic_batch(B, H, W, 4, 1)
        x, y = x.to(dev).contiguous(memory_format=torch.channels_last), y.to(dev)
        net.eval()

        def step():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                p = net(x)
            m = p.argmax(1)
            return torch.bincount((y * 4 + m).reshape(-1), minlength=16)
        ms = timed(step, 4, 2)
        out.update(images_per_s=round(B / (ms / 1e3), 1), ms_per_step=round(ms, 3), batch=B, hw=[H, W])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
