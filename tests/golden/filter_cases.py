"""Crafted candidate lists for the filter_coarse fixture (shared by make_golden.py, which feeds them to the live
reference, and tests/test_oracle_golden.py, which feeds them to the oracle)."""
import torch

FILTER_CASES = [
    # name, rows ('dup' = rows with mutual duplicates, 'nodup' = all distinct), ncn_thres, mutual, ptmax, numpy seed
    ('mutual', 'dup', 0.0, True, None, 0),
    ('mutual_nodup_skips_filter', 'nodup', 0.0, True, None, 0),
    ('nomutual_unique_sorted', 'dup', 0.0, False, None, 0),
    ('thres_partial', 'dup', 0.5, True, None, 0),
    ('thres_empties_is_skipped', 'dup', 2.0, True, None, 0),
    ('ptmax_fill', 'dup', 0.0, True, 50, 11),
    ('ptmax_cut', 'dup', 0.0, True, 5, 12),
    ('ptmax_thres_partial', 'dup', 0.5, True, 9, 13),
    ('ptmax_degenerate_ids', 'dup', 2.0, True, 7, 14),
    ('ptmax_nodup', 'nodup', 0.0, True, 6, 15),
]


def filter_case_inputs(kind):
    """Candidate rows as cal_coarse_matches emits them: dir-1 rows then dir-2 rows, a mutual pair = one row in each."""
    g = torch.Generator().manual_seed(5 if kind == 'dup' else 6)
    d1 = torch.randint(0, 6, (40, 4), generator=g) * 16 + 4
    d1 = torch.unique(d1, dim=0)[torch.randperm(len(torch.unique(d1, dim=0)), generator=g)]
    d2 = torch.randint(6, 12, (len(d1), 4), generator=g) * 16 + 4          # disjoint from d1
    if kind == 'dup':
        d2[::3] = d1[torch.randperm(len(d1), generator=g)[:len(d2[::3])]]  # every third dir-2 row repeats a dir-1 row
    rows = torch.cat([d1, d2], 0)
    if kind == 'nodup':
        rows = torch.unique(rows, dim=0)
        rows = rows[torch.randperm(len(rows), generator=g)]                # distinct rows in a non-sorted order
    scores = torch.rand(len(rows), generator=g)
    return rows, scores
