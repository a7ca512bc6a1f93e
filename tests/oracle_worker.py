"""Worker process of pcsf_helpers._run_workers: runs the CPU oracle (test infrastructure) on a slice of the work.
    python oracle_worker.py in.pkl out.pkl"""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def main():
    from oracle import oracle as o

    job = pickle.load(open(sys.argv[1], "rb"))
    out = []
    if job["kind"] == "mle":
        ps = o.load_paramset(os.path.join(job["params_base"], "PhyloCSF_Parameters", job["pset"]), o.Options(strategy="mle"))
        for codes in job["items"]:
            row = []
            for inst in (ps.model.coding_model, ps.model.noncoding_model):
                tr = {}
                x, (lp, el) = o.maximize_lpr(lambda r: o.lpr_leaves(inst, codes, r), lambda r: r[0], init=1.0, trace=tr)
                row.append((x, lp, el, tr.get("iterations", 0), tr.get("random_tries", 0)))
                inst.q._memo.clear()  # one 32 KB P(t) per branch and candidate would pile up otherwise
            out.append(row)
    elif job["kind"] == "lines":
        opts = o.Options(**job["extra"])
        ps = o.load_paramset(os.path.join(job["params_base"], "PhyloCSF_Parameters", job["pset"]), opts)
        for name, lines in job["items"]:
            out.append(o.process_alignment(ps, opts, name, lines))
    elif job["kind"] == "omega":  # OmegaModel.score (src/OmegaModel.ml:195-219) with the diagnostics unrounded
        import math

        tree = o.load_paramset(os.path.join(job["params_base"], "PhyloCSF_Parameters", job["pset"]), o.Options(strategy="omega")).tree
        omega_H1, sigma_H1 = job["extra"]
        for codes in job["items"]:
            i0 = o.OmegaInstance(tree, [2.5, 1.0, 1.0] + [1.0] * 9, 1.0)
            inst0, l0 = o.omega_kr_map(codes, o.omega_update_f3x4(i0, codes))
            qs = list(inst0.q_settings)
            qs[1], qs[2] = omega_H1, sigma_H1
            inst1, l1 = o.omega_kr_map(codes, inst0.with_q(qs))
            out.append((10.0 * (l1 - l0) / math.log(10.0), l0, inst0.tree_scale, inst0.q_settings[0], l1, inst1.tree_scale, inst1.q_settings[0]))
    else:
        raise SystemExit("unknown job kind")
    pickle.dump(out, open(sys.argv[2], "wb"))


if __name__ == "__main__":
    main()
