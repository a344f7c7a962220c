"""Weight-only-quantised linear at decode sizes: fused dequant+MMA kernel vs bf16 cuBLAS (device-timed, L2 flushed)."""
import json
import sys

import torch

from deepspeed_b200.inference.quantization.layers import maybe_quantized_linear, quantize_weight


def main():
    REP = 6

    def t(fns):
        """fns: callables touching different weights (working set > L2); captured back to back in one CUDA graph so the
        python launch cost is outside the measurement."""
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for f in fns:
                f()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(REP):
                    fns[i % len(fns)]()
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            s, e = torch.cuda.Event(True), torch.cuda.Event(True)
            s.record()
            g.replay()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / REP)
        return sorted(ts)[3]

    rows = []
    for (N, K) in ((28672, 4096), (4096, 14336), (6144, 4096)):
        copies = 3 if N * K > 60e6 else 6
        ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(copies)]
        qws = {m: [quantize_weight(w, m, 128) for w in ws] for m in ("int8", "int4", "fp8")}
        for M in (1, 8, 16):
            x = torch.randn(M, K, device="cuda").bfloat16()
            tb = t([(lambda w=w: torch.nn.functional.linear(x, w)) for w in ws])
            for mode in ("int8", "int4", "fp8"):
                tq = t([(lambda q=q: maybe_quantized_linear(x, q)) for q in qws[mode]])
                by = N * K * (0.5 if mode == "int4" else 1)
                rows.append({"N": N, "K": K, "M": M, "mode": mode, "fused_us": round(tq * 1e3, 1),
                             "weight_stream_GBps": round(by / tq / 1e6), "bf16_cublas_us": round(tb * 1e3, 1),
                             "speedup_vs_bf16": round(tb / tq, 2)})
                print(rows[-1], flush=True)
        del ws, qws
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
