"""CPU emulation of the f16tc rounding points of the HiFi-GAN path with per-stage knobs, to find the
cheapest configuration that holds 1e-3 on the harder "variance preserving" init (TEST/DESIGN TOOL, uses oracle/)."""
import itertools, sys, math
sys.path.insert(0, "oracle"); sys.path.insert(0, "synth_weights")
import torch, torch.nn.functional as F
import gan_oracle as go, ref_init as ri

h = lambda t: t.half().float()
ident = lambda t: t

def forward(sd, cfg, mel, P):
    """P: dict with per-stage lists: res16[i], qa[i], qw[i] (resblock operand rounding), uqa[i], uqw[i] (ups)"""
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], 1, 3)
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        xa = F.leaky_relu(x, 0.1)
        xa = P["uqa"][i](xa)
        x = F.conv_transpose1d(xa, P["uqw"][i](sd[f"ups.{i}.weight"]), sd[f"ups.{i}.bias"], u, u // 2 + u % 2, u % 2)
        qa, qw, r16 = P["qa"][i], P["qw"][i], P["res16"][i]
        def store(x):  # residual stream stored as fp16(lrelu(x)) and recovered by exact inverse
            if not r16: return x
            y = h(F.leaky_relu(x, 0.1))
            return torch.where(y >= 0, y, y * 10)
        x = store(x)
        xs = None
        for j, (kk, dd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            y = x
            pre = f"resblocks.{i*nk+j}"
            for m, d in enumerate(dd):
                xt = qa(F.leaky_relu(y, 0.1))
                xt = F.conv1d(xt, qw(sd[f"{pre}.convs1.{m}.weight"]), sd[f"{pre}.convs1.{m}.bias"], 1, go.get_padding(kk, d), d)
                xt = qa(F.leaky_relu(xt, 0.1))
                xt = F.conv1d(xt, qw(sd[f"{pre}.convs2.{m}.weight"]), sd[f"{pre}.convs2.{m}.bias"], 1, go.get_padding(kk, 1), 1)
                y = xt + y
                if m < len(dd) - 1: y = store(y)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], 1, 3)
    return torch.tanh(x)

def P(res16=(1,1,1,0), qa="hhhh", qw="hhhh", uqa="hhhh", uqw="hhhh"):
    m = {"h": h, "x": ident}
    return dict(res16=res16, qa=[m[c] for c in qa], qw=[m[c] for c in qw], uqa=[m[c] for c in uqa], uqw=[m[c] for c in uqw])

cfg = ri.HIFIGAN_CONFIG_16K
cases = {}
sd0 = go.fold_weight_norm(ri.hifigan_state_dict(cfg, 0))
cases["refinit"] = (sd0, torch.rand(2, 80, 64, generator=torch.Generator().manual_seed(9)) * 8 - 4)
sdr = ri.rescale_variance_preserving(ri.hifigan_state_dict(cfg, 0), 1.0)
cases["rescaled"] = (sdr, torch.rand(2, 80, 64, generator=torch.Generator().manual_seed(9)) * 8 - 4)

variants = {
  "shipped r01 (all fp16, res16 s0-2)": P(),
  "ups split": P(uqa="xxxx", uqw="xxxx"),
  "ups split + s3 split": P(qa="hhhx", qw="hhhx", uqa="xxxx", uqw="xxxx"),
  "ups + s3 split, no res16": P(res16=(0,0,0,0), qa="hhhx", qw="hhhx", uqa="xxxx", uqw="xxxx"),
  "ups + s2,s3 split": P(qa="hhxx", qw="hhxx", uqa="xxxx", uqw="xxxx"),
  "ups + s2,s3 split, res16 s0-1": P(res16=(1,1,0,0), qa="hhxx", qw="hhxx", uqa="xxxx", uqw="xxxx"),
  "ups + s3 split + s2 weights split": P(qa="hhhx", qw="hhxx", uqa="xxxx", uqw="xxxx"),
  "ups + s3 split + s2 act split": P(qa="hhxx", qw="hhhx", uqa="xxxx", uqw="xxxx"),
  "weights split everywhere": P(qw="xxxx", uqw="xxxx"),
  "acts split everywhere, no res16": P(res16=(0,0,0,0), qa="xxxx", uqa="xxxx"),
  "only res16": P(qa="xxxx", qw="xxxx", uqa="xxxx", uqw="xxxx"),
  "s0 only fp16 (+res16 s0)": P(res16=(1,0,0,0), qa="hxxx", qw="hxxx", uqa="xxxx", uqw="xxxx"),
  "s0,s1 fp16 (+res16)": P(res16=(1,1,0,0), qa="hhxx", qw="hhxx", uqa="xxxx", uqw="xxxx"),
}
if __name__ == "__main__":
    sel = sys.argv[1:]
    with torch.no_grad():
        for cname, (sd, mel) in cases.items():
            ref = go.hifigan_forward(sd, cfg, mel)
            print(f"== {cname}: |ref|max {float(ref.abs().max()):.3f} rms {float(ref.pow(2).mean().sqrt()):.3f}")
            for name, p in variants.items():
                if sel and not any(s in name for s in sel): continue
                e = go.rel_errors(forward(sd, cfg, mel, p), ref)
                print(f"  {name:45s} max_rel {e['max_rel']:.2e} rms_rel {e['rms_rel']:.2e}")
