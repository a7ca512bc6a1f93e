"""Synthetic alignment generator for tests and bench.py: columns simulated under a shipped tree and
ECM the way the reference's own test harness does it (PhyloModel.simulate,
lib/CamlPaml/PhyloModel.ml:38-54, with Tools.random_chooser's cumulative-sum inverse sampling,
lib/CamlPaml/Tools.ml:22-41; columns whose reference-species codon is a stop are redrawn,
src/testSim.ml:52-65). torch is used for device memory and its counter-based (Philox) generator;
this is input plumbing, not part of the scoring path."""
import numpy as np
import torch

STOPS = (48, 50, 56)


def parents_from_children(n_leaves, children):
    n = 2 * n_leaves - 1
    par = np.full(n, -1, dtype=np.int64)
    ch = np.asarray(children).reshape(-1, 2)
    for i in range(n_leaves, n):
        par[ch[i - n_leaves, 0]] = i
        par[ch[i - n_leaves, 1]] = i
    return par


def simulate_codes(P, prior, parents, n_leaves, ncols, gen, device, chunk=1 << 19, redraw_ref_stops=True):
    """P: float64 [n_branches, 64, 64] (row = parent state), prior [64]. Returns uint8 [ncols, n_leaves]."""
    P = torch.as_tensor(P, dtype=torch.float64, device=device)
    cum = torch.cumsum(P, dim=2)
    cum_prior = torch.cumsum(torch.as_tensor(prior, dtype=torch.float64, device=device), 0)
    n = 2 * n_leaves - 1
    stops = torch.tensor(STOPS, device=device)
    out = torch.empty((ncols, n_leaves), dtype=torch.uint8, device=device)
    filled = 0
    while filled < ncols:
        m = min(chunk, int((ncols - filled) * 1.12) + 64)
        st = torch.empty((n, m), dtype=torch.int64, device=device)
        u = torch.rand(m, dtype=torch.float64, device=device, generator=gen) * cum_prior[-1]
        st[n - 1] = torch.searchsorted(cum_prior, u).clamp_(max=63)
        for i in range(n - 2, -1, -1):
            cdf = cum[i].index_select(0, st[parents[i]])  # [m, 64]
            u = torch.rand(m, dtype=torch.float64, device=device, generator=gen) * cdf[:, -1]
            st[i] = (cdf < u[:, None]).sum(dim=1).clamp_(max=63)
        leaves = st[:n_leaves].t()
        if redraw_ref_stops:
            leaves = leaves[~torch.isin(leaves[:, 0], stops)]
        take = min(leaves.shape[0], ncols - filled)
        out[filled:filled + take] = leaves[:take].to(torch.uint8)
        filled += take
    return out


_NT = np.frombuffer(b"ACGT", dtype=np.uint8)


def codes_to_nt(codes_frame0, n_align, n_codons):
    """uint8 [n_align*n_codons, n_leaves] frame-0 codon codes -> ASCII nucleotide rows
    uint8 [n_align, n_leaves, 3*n_codons] (the layout pcsf_batch_upload_alignments takes)."""
    dev = codes_frame0.device
    n_leaves = codes_frame0.shape[1]
    c = codes_frame0.view(n_align, n_codons, n_leaves).permute(0, 2, 1).to(torch.int64)  # [a, leaf, codon]
    nt_lut = torch.as_tensor(_NT.copy(), device=dev)
    trip = torch.stack([nt_lut[c // 16], nt_lut[(c // 4) % 4], nt_lut[c % 4]], dim=-1)  # [a, leaf, codon, 3]
    return trip.reshape(n_align, n_leaves, 3 * n_codons).contiguous()
