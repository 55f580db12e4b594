"""In-kernel timeline of a2f_ffn_ln (one launch at the bench shape M=4800): clock64 stamps of the leader CTAs.
    python tools/ffn_timeline.py
Slots: 0 entry, 1 setup done (barriers, TMEM, cluster sync, griddepcontrol.wait), 2 first operands landed,
3..6 MMAs of phase-1 tile t issued, 7 MMAs of the phase-2 tile issued, 8..11 phase-1 tile t stored and published,
12 epilogue enters the LayerNorm tile, 13 LayerNorm stores issued, 14 stores drained, 15 after the final cluster sync."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from a2f_b200 import ops, lib as L

dev = torch.device("cuda:0")
M, N, F = 4800, 768, 3072
g = torch.Generator().manual_seed(1)
x = torch.randn(M, N, generator=g).bfloat16().to(dev)
w1 = (torch.randn(F, N, generator=g) * N ** -0.5).bfloat16().to(dev)
w2 = (torch.randn(N, F, generator=g) * F ** -0.5).bfloat16().to(dev)
b1, b2 = torch.randn(F, generator=g).to(dev), torch.randn(N, generator=g).to(dev)
gamma, beta = torch.rand(N, generator=g).to(dev), torch.randn(N, generator=g).to(dev)
f = torch.empty((M, F), dtype=torch.bfloat16, device=dev)
out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
lib = L.load()
n_cta = 6 * 19
tls = [torch.zeros(2 * n_cta * 16, dtype=torch.int64, device=dev) for _ in range(4)]
for _ in range(3):
    ops.ffn_ln(x, w1, b1, w2, b2, x, gamma, beta, f, out)
for tl_i in tls:                        # four launches back to back, each with its own stamp buffer
    lib.a2f_debug_set_timeline(tl_i.data_ptr())
    ops.ffn_ln(x, w1, b1, w2, b2, x, gamma, beta, f, out)
lib.a2f_debug_set_timeline(None)
torch.cuda.synchronize()
gt = [tl_i.view(2, n_cta, 16)[1].cpu().double() for tl_i in tls]     # globaltimer (ns)
t0 = gt[0][:, 0].min()
print("launch   first entry   median entry   last entry | median setup-done | first exit   median exit   last exit   (us, globaltimer)")
for i, g_ in enumerate(gt):
    en, sd, ex = (g_[:, 0] - t0) / 1e3, (g_[:, 1] - t0) / 1e3, (g_[:, 15] - t0) / 1e3
    print(f"{i:6d} {en.min():13.2f} {en.median():14.2f} {en.max():12.2f} | {sd.median():17.2f} | {ex.min():10.2f} {ex.median():13.2f} {ex.max():11.2f}")
t = tls[1].view(2, n_cta, 16)[0].cpu().double()
t = (t - t[:, :1])                      # cycles since this CTA's entry
lead = t[0::2]                          # leader CTAs (MMA stamps live there)
names = ["entry", "setup done", "first operands", "mma t0", "mma t1", "mma t2", "mma t3", "mma phase2", "pub t0", "pub t1",
         "pub t2", "pub t3", "enter LN", "LN stores issued", "stores drained", "after cluster sync"]
clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1900
print(f"cycles since CTA entry, leaders of {lead.shape[0]} pairs (median / min / max); ~{clk} MHz")
for i, nme in enumerate(names):
    col = lead[:, i]
    print(f"{i:2d} {nme:18s} {col.median():9.0f} {col.min():9.0f} {col.max():9.0f}   {col.median() / clk:7.2f} us")
