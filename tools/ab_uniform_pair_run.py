"""One timed pass of the uniform sampler's whole `a` chain (shared by the A/B scripts)."""
import torch


def run(ctx, stream, n, np_, batch):
    gen = torch.Generator(device="cuda").manual_seed(3)
    d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((batch, np_, n), dtype=torch.int32, device="cuda")
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    def step():
        d_ctr.zero_()
        for p in range(np_):
            ctx.sample_uniform_device(d_ss, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3): step()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3, int(d_out.view(-1)[::1031].to(torch.int64).sum().item()) ^ int(d_ctr.to(torch.int64).sum().item())
