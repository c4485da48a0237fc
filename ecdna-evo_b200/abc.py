"""ABC record writer: the CSV of the reference's removed `abc` binary (abc.md:38-55), one row per
prior draw, ALL draws saved so the thresholds can be tuned afterwards (abc.md:57-71).

Input is what the kernel's fused epilogue returns per draw (distances, accept flag, final counts) and
the prior draws; nothing is simulated here."""
import csv

ABC_FIELDS = ["parental_idx", "idx", "timepoint", "seed", "ecdna", "mean", "entropy", "f1", "f2", "d1", "d2", "cells",
              "tumour_cells", "init_mean", "init_cells", "init_copies"]


def abc_rows(opts, idx_begin, rates, abc_distance, nminus, nplus, timepoint=0, sample_cells=None, parental_idx=None):
    """Yield one dict per draw with the columns of abc.md:38-55.

    rates[i] = (b0, b1, d0, d1).  abc.md's naming: f1/d1 belong to the cells WITH ecDNA, f2/d2 to the
    cells WITHOUT.  abc_distance[i] = (ks, rel.mean, rel.entropy, rel.frequency)."""
    init_cells = sum(opts.distribution.values())
    init_copies = sum(k * c for k, c in opts.distribution.items())
    init_mean = init_copies / init_cells if init_cells else 0.0
    for i in range(len(rates)):
        tumour = int(nminus[i]) + int(nplus[i])
        yield {
            "parental_idx": "" if parental_idx is None else parental_idx, "idx": idx_begin + i, "timepoint": timepoint,
            "seed": opts.seed, "ecdna": float(abc_distance[i][0]), "mean": float(abc_distance[i][1]),
            "entropy": float(abc_distance[i][2]), "f1": float(rates[i][1]), "f2": float(rates[i][0]),
            "d1": float(rates[i][3]), "d2": float(rates[i][2]),
            "cells": tumour if sample_cells is None else int(sample_cells), "tumour_cells": tumour,
            "init_mean": init_mean, "init_cells": init_cells, "init_copies": init_copies,
        }


def write_abc_csv(path, opts, idx_begin, rates, abc_distance, nminus, nplus, **kw):
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=ABC_FIELDS)
        w.writeheader()
        n = 0
        for row in abc_rows(opts, idx_begin, rates, abc_distance, nminus, nplus, **kw):
            w.writerow(row)
            n += 1
    return n
