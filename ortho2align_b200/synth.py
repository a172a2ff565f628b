"""Seeded synthetic HSP/background data of the shapes BASELINE.json names.

Null scores:  max(12, rint(Gumbel(17, 3.2)))  (BLASTN word_size 6 => min raw score 12)
Signal:       3 % of HSPs get + rint(Exp(18)) and are laid out colinearly (a true ortholog)
Cut HSPs:     2 % of HSPs get score *= U(0.3, 1)  (float scores, `alignment.py:330`)
Coordinates:  query gene window (from a BED6 or a U-placed 20 kb window); per region a 2.1 Mb
              subject window; HSP length U[12, 120]; 50 % reverse orientation.
Background:   the same null, `n_bg` scores per lncRNA.

Everything is generated directly in the columnar layout (`columnar.ColumnarBatch`);
`to_json_dicts` renders a (small) batch in the reference's JSON schema
(`genomicranges.py:1077-1082`, `alignment.py:588-601`, `genomicranges.py:796-814`).
The unit of seeding is a block of lncRNAs: `default_rng([seed, config_id, block_index])`, so any
rank can regenerate exactly its shard.
"""
from __future__ import annotations

import os

import numpy as np

from .columnar import ColumnarBatch, concat_batches

SUBJECT_WINDOW = 2_100_000
GENE_WINDOW = 20_000
_STRRNA_GENES = os.path.join(os.path.dirname(__file__), "data", "strrna_genes.json")


def null_scores(rng, n):
    s = np.rint(rng.gumbel(17.0, 3.2, size=n))
    return np.maximum(12.0, s)


def _gen_block(rng, genes, hsp_counts, aln_off, n_bg, dense=False):
    """genes: int64 [n_lnc, 2] (start, end); hsp_counts: int64 [n_aln]; aln_off [n_lnc+1]."""
    n_lnc = len(genes)
    n_aln = len(hsp_counts)
    hsp_off = np.zeros(n_aln + 1, dtype=np.int64)
    np.cumsum(hsp_counts, out=hsp_off[1:])
    n = int(hsp_off[-1])
    aln_of_hsp = np.repeat(np.arange(n_aln), hsp_counts)
    lnc_of_aln = np.repeat(np.arange(n_lnc), np.diff(aln_off))
    gstart = genes[lnc_of_aln, 0][aln_of_hsp]
    glen = (genes[lnc_of_aln, 1] - genes[lnc_of_aln, 0])[aln_of_hsp]
    # subject window per alignment
    s0_aln = rng.integers(1_000_000, 200_000_000, size=n_aln)
    s0 = s0_aln[aln_of_hsp]
    length = rng.integers(12, 121, size=n)
    length = np.minimum(length, np.maximum(glen - 1, 1))
    qpos = (rng.random(n) * np.maximum(glen - length, 1)).astype(np.int64)
    spos = (rng.random(n) * (SUBJECT_WINDOW - 121)).astype(np.int64)
    slength = np.maximum(1, length + rng.integers(-3, 4, size=n))
    reverse = rng.random(n) < 0.5
    score = null_scores(rng, n)
    if dense:
        # C4: every HSP scores above any background value -> p = 0, all enter the DP
        score = score + 200.0
    else:
        signal = rng.random(n) < 0.03
        score = score + np.where(signal, np.rint(rng.exponential(18.0, size=n)), 0.0)
        # colinear placement of the signal HSPs of an alignment around a diagonal whose
        # orientation is fixed per alignment
        aln_rev = (rng.random(n_aln) < 0.5)[aln_of_hsp]
        diag0 = (rng.random(n_aln) * (SUBJECT_WINDOW - 2 * GENE_WINDOW - 1000)).astype(np.int64)[aln_of_hsp] + 500
        jitter = np.rint(rng.normal(0.0, 40.0, size=n)).astype(np.int64)
        sp_dir = diag0 + qpos + jitter
        sp_rev = diag0 + (glen - qpos) + jitter
        sp_sig = np.clip(np.where(aln_rev, sp_rev, sp_dir), 0, SUBJECT_WINDOW - 130)
        spos = np.where(signal, sp_sig, spos)
        reverse = np.where(signal, aln_rev, reverse)
        cut = rng.random(n) < 0.02
        score = np.where(cut, score * rng.uniform(0.3, 1.0, size=n), score)
    qstart = gstart + qpos
    qend = qstart + length
    lo = s0 + spos
    hi = lo + slength
    sstart = np.where(reverse, hi, lo)
    send = np.where(reverse, lo, hi)
    # qlen/slen as cut_coordinates recomputes them: max+1 over all HSPs (alignment.py:477-487)
    qlen = np.ones(n_aln, dtype=np.int64)
    slen = np.ones(n_aln, dtype=np.int64)
    nz = hsp_counts > 0
    if n:
        starts = hsp_off[:-1][nz]
        qlen[nz] = np.maximum.reduceat(qend, starts) + 1
        slen[nz] = np.maximum.reduceat(np.maximum(sstart, send), starts) + 1
    bg = null_scores(rng, n_lnc * n_bg).astype(np.int32)
    bg_off = np.arange(n_lnc + 1, dtype=np.int64) * n_bg
    meta = {"s0": s0_aln}
    return ColumnarBatch(
        bg_off=bg_off, bg=bg, aln_off=aln_off.astype(np.int64), hsp_off=hsp_off, qlen=qlen, slen=slen,
        qstart=qstart.astype(np.int32), qend=qend.astype(np.int32),
        sstart=sstart.astype(np.int32), send=send.astype(np.int32),
        score=score.astype(np.float64)), meta


def load_strrna_bed(path=_STRRNA_GENES):
    """The 96 genes (chrom, start, end, name, strand) of BASELINE.json config 1."""
    import json
    with open(path) as f:
        return [tuple(g) for g in json.load(f)["genes"]]


def workload(config, n_lnc=None, seed=0, lnc_range=None, block=64, with_meta=False,
             n_bg=None, hsps_per_region=None, regions=None):
    """Generate the lncRNAs [lnc_range) of BASELINE.json config `config` (1..5).

    Returns ColumnarBatch (names filled) and, if `with_meta`, a list of per-lncRNA dicts with the
    gene record and per-alignment subject windows (used by `to_json_dicts`).
    """
    cfg = int(config)
    bed = None
    if cfg == 1:
        bed = load_strrna_bed()
        total = len(bed) if n_lnc is None else n_lnc
        block = 1
    elif cfg in (2, 3):
        total = 20_000 if n_lnc is None else n_lnc
    elif cfg == 4:
        total = 5_000 if n_lnc is None else n_lnc
    elif cfg == 5:
        total = 100_000 if n_lnc is None else n_lnc
    else:
        raise ValueError("config must be 1..5")
    lo, hi = (0, total) if lnc_range is None else lnc_range
    parts, metas = [], []
    b0 = (lo // block) * block
    for bstart in range(b0, hi, block):
        bend = min(bstart + block, total)
        nb = bend - bstart
        rng = np.random.default_rng([seed, cfg, bstart // block])
        if cfg == 1:
            g = bed[bstart % len(bed)]
            genes = np.array([[g[1], g[2]]], dtype=np.int64)
            n_reg = rng.integers(1, 6, size=nb) if regions is None else np.full(nb, regions)
            aln_off = np.concatenate([[0], np.cumsum(n_reg)])
            counts = rng.integers(200, 2001, size=int(aln_off[-1])) if hsps_per_region is None \
                else np.full(int(aln_off[-1]), hsps_per_region)
            nbg = 1000 if n_bg is None else n_bg
            dense = False
        elif cfg in (2, 3):
            gs = rng.integers(1_000_000, 200_000_000, size=nb)
            genes = np.stack([gs, gs + GENE_WINDOW], axis=1)
            r = 5 if regions is None else regions
            aln_off = np.arange(nb + 1) * r
            h = 2000 if hsps_per_region is None else hsps_per_region
            counts = rng.integers(int(h * 0.9), int(h * 1.1) + 1, size=nb * r)
            nbg = (10_000 if cfg == 2 else 1_000_000) if n_bg is None else n_bg
            dense = False
        elif cfg == 4:
            gs = rng.integers(1_000_000, 200_000_000, size=nb)
            genes = np.stack([gs, gs + GENE_WINDOW], axis=1)
            r = 1 if regions is None else regions
            aln_off = np.arange(nb + 1) * r
            h = 50_000 if hsps_per_region is None else hsps_per_region
            counts = np.full(nb * r, h)
            nbg = 1000 if n_bg is None else n_bg
            dense = True
        else:
            gs = rng.integers(1_000_000, 200_000_000, size=nb)
            glen = rng.integers(300, GENE_WINDOW, size=nb)
            genes = np.stack([gs, gs + glen], axis=1)
            n_reg = 1 + rng.poisson(1.0, size=nb) if regions is None else np.full(nb, regions)
            aln_off = np.concatenate([[0], np.cumsum(n_reg)])
            mean_h = 100.0 if hsps_per_region is None else float(hsps_per_region)
            sigma = 1.0
            counts = np.maximum(1, rng.lognormal(np.log(mean_h) - sigma * sigma / 2, sigma,
                                                 size=int(aln_off[-1])).astype(np.int64))
            nbg = 1000 if n_bg is None else n_bg
            dense = False
        part, meta = _gen_block(rng, genes, np.asarray(counts, dtype=np.int64),
                                np.asarray(aln_off, dtype=np.int64), nbg, dense=dense)
        if cfg == 1:
            part.names = [bed[bstart % len(bed)][3]]
            chroms = [bed[bstart % len(bed)][0]]
            strands = [bed[bstart % len(bed)][4]]
        else:
            part.names = [f"lnc{cfg}_{i}" for i in range(bstart, bend)]
            chroms = ["chr1"] * nb
            strands = ["+" if (i % 2 == 0) else "-" for i in range(bstart, bend)]
        # trim to requested range inside the block
        s_lo, s_hi = max(lo, bstart) - bstart, min(hi, bend) - bstart
        if with_meta:
            for i in range(s_lo, s_hi):
                a0, a1 = int(part.aln_off[i]), int(part.aln_off[i + 1])
                metas.append({"chrom": chroms[i], "start": int(genes[i, 0]), "end": int(genes[i, 1]),
                              "strand": strands[i], "name": part.names[i],
                              "s0": [int(x) for x in meta["s0"][a0:a1]]})
        if s_lo != 0 or s_hi != nb:
            part = part.slice_lnc(s_lo, s_hi)
        parts.append(part)
    batch = concat_batches(parts)
    if with_meta:
        return batch, metas
    return batch


def _grange_dict(chrom, start, end, strand, name, relations=None, relations_class=None):
    return {"chrom": chrom, "start": start, "end": end, "strand": strand, "genome": None,
            "sequence_file_path": {"path": None}, "name": name,
            "relations": {} if relations is None else relations,
            "relations_class": relations_class}


def to_json_dicts(batch, metas):
    """Render each lncRNA of a (small) batch in the reference JSON schema.

    Returns list of (name, alignment_dict_list, background_list)."""
    out = []
    for l in range(batch.n_lnc):
        m = metas[l]
        a0, a1 = int(batch.aln_off[l]), int(batch.aln_off[l + 1])
        sranges = []
        for k, a in enumerate(range(a0, a1)):
            s0 = m["s0"][k]
            sranges.append(("chr" + str(1 + (s0 % 19)), s0, s0 + SUBJECT_WINDOW))
        lifted = {"collection": [_grange_dict(c, s + 50_000, e - 50_000, ".", m["name"])
                                 for (c, s, e) in sranges],
                  "sequence_file_path": {"path": None}, "grange_class": "GenomicRange"}
        alns = []
        for k, a in enumerate(range(a0, a1)):
            h0, h1 = int(batch.hsp_off[a]), int(batch.hsp_off[a + 1])
            hsps = []
            for h in range(h0, h1):
                sc = float(batch.score[h])
                hsps.append({"qstart": int(batch.qstart[h]), "qend": int(batch.qend[h]),
                             "sstart": int(batch.sstart[h]), "send": int(batch.send[h]),
                             "score": int(sc) if sc.is_integer() else sc,
                             "pval": None, "kwargs": {}})
            c, s, e = sranges[k]
            alns.append({"HSPs": hsps, "qlen": int(batch.qlen[a]), "slen": int(batch.slen[a]),
                         "filtered_HSPs": [dict(h) for h in hsps], "filtered": False,
                         "qrange": _grange_dict(m["chrom"], m["start"], m["end"], m["strand"], m["name"],
                                                relations={"lifted": lifted},
                                                relations_class="GenomicRangesList"),
                         "srange": _grange_dict(c, s, e, ".", None),
                         "relative": False})
        b0, b1 = int(batch.bg_off[l]), int(batch.bg_off[l + 1])
        out.append((m["name"], alns, [int(x) for x in batch.bg[b0:b1]]))
    return out


# --------------------------------------------------------------------------------------------------
# Producer side: what `_align_syntenies` (pipeline.py:507-546) sees before it
# writes the alignment JSON — BLAST "7 std score" tables of a flanked query neighbourhood against a
# syntenic window, in sequence-relative 1-based inclusive coordinates.
# --------------------------------------------------------------------------------------------------
BLAST_FIELDS = ("query acc.ver, subject acc.ver, % identity, alignment length, mismatches, gap opens, "
                "q. start, q. end, s. start, s. end, evalue, bit score, score")


def relative_hits(rng, n, q_len, s_len):
    """n BLAST hits as 1-based inclusive (qstart, qend, sstart, send, score) rows; 50 % minus-strand subject
    (s. start > s. end, as blastn prints them)."""
    ln = rng.integers(12, 121, size=n)
    qs = rng.integers(1, max(2, q_len - 121), size=n)
    qe = qs + ln - 1
    ss = rng.integers(1, max(2, s_len - 121), size=n)
    se = ss + ln - 1 + rng.integers(-3, 4, size=n)          # gapped: subject span differs a little
    rev = rng.random(n) < 0.5
    ss, se = np.where(rev, se, ss), np.where(rev, ss, se)
    return qs, qe, ss, se, null_scores(rng, n).astype(np.int64)


def blast_tabular(qs, qe, ss, se, score, qname="query", sname="subject", fields=BLAST_FIELDS):
    """Render hits as `blastn -outfmt "7 std score"` text (the stream `from_file_blast` parses,
    alignment.py:515-586)."""
    head = [f"# BLASTN 2.12.0+", f"# Query: {qname}", f"# Database: User specified sequence set (Input: {sname}.fa)"]
    if len(qs) == 0:
        return "\n".join(head + ["# 0 hits found", "# BLAST processed 1 queries", ""])
    rows = [f"{qname}\t{sname}\t{90 + (int(s) % 10)}.000\t{abs(int(b) - int(a)) + 1}\t0\t0\t{a}\t{b}\t{c}\t{d}\t1e-05\t{int(s) * 1.8:.1f}\t{s}"
            for a, b, c, d, s in zip(qs, qe, ss, se, score)]
    return "\n".join(head + [f"# Fields: {fields}", f"# {len(qs)} hits found"] + rows + ["# BLAST processed 1 queries", ""])


def producer_workload(n_lnc, seed=0, regions=5, hits_per_region=2000, flank=50_000, strands=("+", "-", ".")):
    """Per query gene: a flanked neighbourhood aligned to `regions` syntenic windows.

    Returns (rel, ranges): `rel` = dict of flat arrays in the layout `o2a_cut_batch` takes (hsp_off and
    0-based half-open relative coordinates as `from_file_blast` leaves them, float64 scores); `ranges` =
    per-alignment int64 arrays q_start, q_end, q_minus, s_start, s_end, s_minus (the neighbourhood and the
    syntenic window, genomicranges.py:1108-1129) and qleft, qright (the gene itself, pipeline.py:530-531)."""
    rng = np.random.default_rng([seed, 77])
    n_aln = n_lnc * regions
    counts = np.maximum(1, rng.poisson(hits_per_region, size=n_aln)) if hits_per_region > 0 else np.zeros(n_aln, np.int64)
    hsp_off = np.zeros(n_aln + 1, dtype=np.int64)
    np.cumsum(counts, out=hsp_off[1:])
    n = int(hsp_off[-1])
    aln_of = np.repeat(np.arange(n_aln), counts)
    gene_len = np.repeat(rng.integers(300, GENE_WINDOW, size=n_lnc), regions)
    gene_start = np.repeat(rng.integers(flank + 1, 200_000_000, size=n_lnc), regions)
    q_start, q_end = gene_start - flank, gene_start + gene_len + flank
    s_start = rng.integers(1, 200_000_000, size=n_aln)
    s_end = s_start + SUBJECT_WINDOW
    q_minus = np.repeat(rng.choice([s == "-" for s in strands], size=n_lnc), regions).astype(np.int8)
    s_minus = rng.choice([s == "-" for s in strands], size=n_aln).astype(np.int8)
    # hits concentrate around the gene (60 %) so that a realistic share survives the cut and some straddle it
    q_span = (q_end - q_start)[aln_of]
    near = rng.random(n) < 0.6
    centre = np.where(q_minus[aln_of] == 1, q_span - flank - gene_len[aln_of] // 2, flank + gene_len[aln_of] // 2)
    qs = np.where(near, centre + (rng.normal(0, 1, size=n) * gene_len[aln_of]).astype(np.int64),
                  rng.integers(0, 1 << 62, size=n) % np.maximum(1, q_span - 130))
    ln = rng.integers(12, 121, size=n)
    qs = np.clip(qs, 0, q_span - 130)
    qe = qs + ln
    ss = rng.integers(0, SUBJECT_WINDOW - 130, size=n)
    se = ss + ln + rng.integers(-3, 4, size=n)
    rev = rng.random(n) < 0.5
    # minus-strand subject hit as from_file_blast leaves it: (s. start - 1, s. end) with s. start > s. end
    ss, se = np.where(rev, se - 1, ss), np.where(rev, ss + 1, se)
    rel = dict(hsp_off=hsp_off, qstart=qs.astype(np.int32), qend=qe.astype(np.int32), sstart=ss.astype(np.int32),
               send=se.astype(np.int32), score=null_scores(rng, n).astype(np.float64))
    ranges = dict(q_start=q_start.astype(np.int64), q_end=q_end.astype(np.int64), q_minus=q_minus,
                  s_start=s_start.astype(np.int64), s_end=s_end.astype(np.int64), s_minus=s_minus,
                  qleft=gene_start.astype(np.int64), qright=(gene_start + gene_len).astype(np.int64))
    return rel, ranges


# ---------------------------------------------------------------------------------------------
# annotate_orthologs: orthologs and a gene annotation as the columns of o2a_ann_input
# ---------------------------------------------------------------------------------------------
def annotate_workload(n_q, n_a, seed=0, n_chrom=24, span=100_000_000, inverted=0.0, long_frac=0.01):
    """-> (q, a): dicts of chrom, start, end, name (int32) and strand (int8: +1, -1, 0), both sorted by the
    reference's list key (chrom, start, end, name). Annotation lengths are gene-like (log-normal around 8 kb,
    `long_frac` of them 0.2-2 Mb — the ranges that make a backward scan long); orthologs are 0.2-6 kb chains.
    `inverted` = fraction of ranges with end < start (the reference's loops accept them)."""
    rng = np.random.default_rng(4000 + seed)

    def make(n, lengths):
        chrom = rng.integers(0, n_chrom, n).astype(np.int32)
        start = rng.integers(0, max(span, 1), n).astype(np.int64)
        end = start + lengths
        if inverted > 0 and n:
            flip = rng.random(n) < inverted
            start[flip], end[flip] = end[flip], start[flip].copy()
        name = rng.integers(0, max(n // 3, 1), n).astype(np.int32)
        strand = rng.choice(np.array([1, -1, 0], dtype=np.int8), n, p=[0.45, 0.45, 0.1])
        order = np.lexsort((name, end, start, chrom))
        return dict(chrom=chrom[order], start=np.minimum(start[order], 2**31 - 1).astype(np.int32),
                    end=np.minimum(end[order], 2**31 - 1).astype(np.int32), name=name[order], strand=strand[order])

    a_len = np.minimum(rng.lognormal(9.0, 1.2, n_a), 5e5).astype(np.int64)
    if n_a:
        long_ones = rng.random(n_a) < long_frac
        a_len[long_ones] = rng.integers(200_000, 2_000_000, int(long_ones.sum()))
        a_len[rng.random(n_a) < 0.02] = 0                          # zero-length ranges (touching cases)
    q_len = rng.integers(200, 6000, n_q).astype(np.int64)
    return make(n_q, q_len), make(n_a, a_len)
