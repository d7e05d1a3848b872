import sys
import torch
a, b = torch.load(sys.argv[1]), torch.load(sys.argv[2])
B = 32
for k in ("lin", "stop", "dlin", "dstop", "loss", "dmem", "g0pre", "m1", "act0", "dG1", "dG0"):
    if k in a and k in b:
        x, y = a[k].double(), b[k].double()
        e = (x - y).abs()
        msg = "%-6s max|a| %.3e max diff %.3e" % (k, x.abs().max().item(), e.max().item())
        if e.dim() == 2 and e.shape[0] % B == 0 and e.shape[0] > B:
            per_t = e.view(-1, B, e.shape[1]).amax(dim=(1, 2))
            top = per_t.topk(5)
            msg += " | worst steps %s diffs %s" % (top.indices.tolist(), ["%.2e" % v for v in top.values.tolist()])
            ref_t = x.abs().view(-1, B, e.shape[1]).amax(dim=(1, 2))
            msg += " | rel at worst %.2e" % (top.values[0] / ref_t[top.indices[0]]).item()
        print(msg)
for k in a["grads"]:
    x, y = a["grads"][k].double(), b["grads"][k].double()
    print("%-26s max|a| %.3e diff %.3e rel %.2e" % (k, x.abs().max().item(), (x - y).abs().max().item(), (x - y).abs().max().item() / x.abs().max().item()))
