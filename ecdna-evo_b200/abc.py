"""ABC record writer: the CSV of the reference's removed `abc` binary (abc.md:38-55), one row per
prior draw, ALL draws saved so the thresholds can be tuned afterwards (abc.md:57-71).

Input is what the kernel's fused epilogue returns per draw (distances, accept flag, final counts) and
the prior draws; nothing is simulated here."""
import csv

import numpy as np

REC_HEADER = 16  # ECDNA_B200_ABC_REC_HEADER: 32-bit words in front of a record's distribution


def record_words(rec_bins):
    return REC_HEADER + rec_bins


def decode_records(words, rec_bins):
    """Records of ecdna_b200_abc_pack ([n][16 + rec_bins] u32) as a dict of arrays (layout: include/ecdna_b200.h)."""
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, record_words(rec_bins))
    f = w.view(np.float32)
    return {"idx": w[:, 0].astype(np.uint64) | (w[:, 1].astype(np.uint64) << np.uint64(32)), "rates": f[:, 2:6].copy(),
            "distance": f[:, 6:10].copy(), "mean": f[:, 10].copy(), "frequency": f[:, 11].copy(),
            "entropy": f[:, 12].copy(), "cells": w[:, 13].copy(), "kmax": w[:, 14].copy(), "stop": w[:, 15] & 0xFF,
            "hist": w[:, REC_HEADER:].copy()}


def merge_gathered(all_records, all_counts, capacity, rec_bins):
    """What ecdna_b200_abc_allgather leaves on every rank - [world][capacity] record blocks and [world] counts -
    as one [n_accepted_total][words] array in rank order (= index order, ranks own contiguous index blocks)."""
    all_counts = np.asarray(all_counts, dtype=np.int64)
    if np.any(all_counts > capacity):
        raise OverflowError(f"a rank accepted {int(all_counts.max())} draws but the record blocks hold {capacity}: "
                            "raise the capacity and run the exchange again")
    blocks = np.asarray(all_records, dtype=np.uint32).reshape(len(all_counts), capacity, record_words(rec_bins))
    return np.concatenate([blocks[r, : all_counts[r]] for r in range(len(all_counts))], axis=0)


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
