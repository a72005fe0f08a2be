#!/usr/bin/env python
"""CPU simulation of the operand splits considered for the fused backward chain: a 4-layer gated data-gradient chain and
one weight gradient computed as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp32 accumulation, against fp64.

    python scripts/sim_split_accuracy.py

Round-1 result (rows spanning ~8 orders of magnitude): bf16 hi/lo pair 5e-6 .. 9e-6 of the tensor's max after 1 .. 4 layers and 9e-6 for
the weight gradient; truncating TF32 pair 5e-7 .. 1.3e-6 -- both inside the gradient tolerance of the parity tests (1e-4 x max)."""
import torch


def split_bf16(x):
    hi = x.to(torch.bfloat16).float()
    return hi, (x - hi).to(torch.bfloat16).float()


def split_tf32(x):                      # what kind::tf32 does with raw fp32 words: the low 13 mantissa bits are ignored
    hi = (x.view(torch.int32) & -8192).view(torch.float32)
    lo = x - hi
    return hi, (lo.view(torch.int32) & -8192).view(torch.float32)


def mm3(a, w, split):
    ah, al = split(a)
    wh, wl = split(w)
    return ah @ wh.T + al @ wh.T + ah @ wl.T


def main():
    torch.manual_seed(0)
    M, N = 4096, 256
    Ws = [torch.randn(N, N) * (2.0 / N) ** 0.5 for _ in range(4)]
    Ys = [torch.randn(M, N) for _ in range(4)]
    d0 = torch.randn(M, N) * torch.exp(torch.randn(M, 1) * 4) * 1e-6
    X = torch.randn(M, N)
    for name, split in (("bf16 hi/lo", split_bf16), ("tf32 hi/lo (truncating)", split_tf32)):
        d, d64 = d0.clone(), d0.double()
        for l in range(4):
            gate = torch.where(Ys[l] > 0, 1.0, 0.01)
            d = mm3(d * gate, Ws[l].T.contiguous(), split)
            d64 = (d64 * gate.double()) @ Ws[l].double()
            print(f"{name}: data gradient after layer {l + 1}: max err / max |ref| = {(d.double() - d64).abs().max() / d64.abs().max():.2e}")
        dW = mm3(d0.T.contiguous(), X.T.contiguous(), split)
        ref = d0.double().T @ X.double()
        print(f"{name}: weight gradient: max err / max |ref| = {(dW.double() - ref).abs().max() / ref.abs().max():.2e}")


if __name__ == "__main__":
    main()
