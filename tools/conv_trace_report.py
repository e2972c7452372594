"""Reads the clock64() handshake trace CTA 0 of the conv backward kernels wrote (EEGB200_CONV_TRACE=<dir>, csrc/conv_tc.cu)
and prints, per role, where the cycles of a steady-state iteration go.     python tools/conv_trace_report.py <dir>"""
import struct
import sys

TRACE_IT, ROLES = 96, 8
NAMES = {0: "builder g0", 1: "builder g1", 2: "epilogue g0", 3: "epilogue g1", 4: "control A", 5: "control B", 6: "scatter"}
EV = {
    "builder": ["start", "pooled(bar)", "im_empty ok", "im2col written+arrive", "dym_empty ok", "dym written"],
    "epilogue": ["start", "c_full ok", "op_empty ok", "done+arrive", "math done (B2)"],
    "control A": ["start", "im_full ok", "c_empty ok", "issued+commit"],
    "control B": ["start", "dym_full ok (B1) / op_full ok (B2)", "op_full ok (B1) / gg_empty ok (B2)", "issued+commit"],
    "scatter": ["start", "gg_full ok", "row stored"],
}


def main(d):
    for kern in ("fwd_stats", "fwd_apply", "bwd_stats", "bwd_apply"):
        try:
            raw = open(f"{d}/conv_trace_{kern}.bin", "rb").read()
        except FileNotFoundError:
            continue
        v = struct.unpack(f"{len(raw) // 8}q", raw)
        t = lambda role, it, ev: v[(role * TRACE_IT + it) * 8 + ev]
        print(f"== {kern} (cycles; iterations 16..79 of CTA 0)")
        for role, name in NAMES.items():
            its = [it for it in range(16, 80) if t(role, it, 0) > 0]
            if len(its) < 4:
                continue
            step = its[1] - its[0]
            period = (t(role, its[-1], 0) - t(role, its[0], 0)) / (its[-1] - its[0])
            kind = name.split(" g")[0]
            print(f"  {name}: {len(its)} iterations traced, {period:.0f} cycles per iteration of the kernel "
                  f"({period * step:.0f} between this role's own iterations)")
            evs = EV[kind]
            for e in range(1, len(evs)):
                ds = [t(role, it, e) - max(t(role, it, k) for k in range(e) if t(role, it, k) > 0) for it in its
                      if t(role, it, e) > 0]
                if ds:
                    print(f"      -> {evs[e]:42s} +{sum(ds) / len(ds):7.0f}  (max {max(ds)}, n={len(ds)})")
        # pipeline lag: control A issue of it -> epilogue sees c_full -> epilogue done -> control B issue
        lag1 = [t(2 + (it & 1), it, 1) - t(4, it, 3) for it in range(16, 80) if t(4, it, 3) and t(2 + (it & 1), it, 1)]
        lag2 = [t(5, it, 3) - t(2 + (it & 1), it, 3) for it in range(16, 80) if t(5, it, 3) and t(2 + (it & 1), it, 3)]
        if lag1:
            print(f"  control A commit -> epilogue sees c_full : {sum(lag1) / len(lag1):.0f} cycles (UMMA completion)")
        if lag2:
            print(f"  epilogue done -> control B issued+commit  : {sum(lag2) / len(lag2):.0f} cycles")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
