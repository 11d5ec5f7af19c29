"""A recording stand-in for libtqdne_b200.so so that the host-side lowering (plan construction, slice tables,
buffer reuse, FLOP accounting) can be exercised on a machine without a GPU.  It computes nothing."""
import ctypes as C


class FakeLib:
    def __init__(self):
        self.calls = []
        self.convs = []
        self.gns = []
        self.attns = []
        self.resamples = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append(name)
            if name == "tq_plan_create":
                return 1
            if name == "tq_conv_stats_parts":
                return 3
            if name == "tq_plan_num_ops":
                return sum(1 for c in self.calls if c.startswith("tq_plan_add_"))
            if name == "tq_plan_add_conv":
                d = args[1]._obj if hasattr(args[1], "_obj") else args[1].contents
                n = d.num_classes * d.num_slices
                self.convs.append(dict(
                    N=d.N, H=d.H, W=d.W, cout=d.cout, cout_pad=d.cout_pad, ktot=d.ktot, num_srcs=d.num_srcs,
                    num_classes=d.num_classes, num_slices=d.num_slices,
                    slices=[(d.slices[i].src, d.slices[i].dx, d.slices[i].dy, d.slices[i].c0, d.slices[i].kb) for i in range(n)],
                    srcs=[(d.srcs[i].N, d.srcs[i].H, d.srcs[i].W, d.srcs[i].C, d.srcs[i].sn, d.srcs[i].sy, d.srcs[i].sx)
                          for i in range(d.num_srcs)],
                    out_strides=(d.out_sn, d.out_sy, d.out_sx), class_off=list(d.out_class_off), out_dtype=d.out_dtype,
                    has_emb=bool(d.emb), has_res=bool(d.residual), stats=d.stats, stats_parts=d.stats_parts))
            if name == "tq_plan_add_groupnorm":
                d = args[1]._obj if hasattr(args[1], "_obj") else args[1].contents
                self.gns.append(dict(N=d.N, P=d.P, C0=d.C0, C1=d.C1, stats0=d.stats0, stats1=d.stats1, ws=d.ws, parts0=d.parts0, parts1=d.parts1,
                                     film=d.film, film_ld=d.film_ld))
            if name == "tq_plan_add_attention":
                d = args[1]._obj if hasattr(args[1], "_obj") else args[1].contents
                self.attns.append(dict(N=d.N, T=d.T, heads=d.heads, d=d.d, causal=d.causal, lse=d.lse))
            if name == "tq_plan_add_resample2":
                self.resamples.append(dict(N=args[4], H=args[5], W=args[6], C=args[7], mode=args[8]))
            if name == "tq_last_error":
                return b""
            return 0
        return fn
