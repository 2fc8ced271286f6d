"""Stand-in for the `pylev` package (not installed in this image, no network) so that the UNMODIFIED reference
`ocrs_models/train_rec.py` (its line 5 imports it, line 64 calls `pylev.levenshtein`) can be imported and run by
the reference arm of bench.py and by the drop-in tests. Same contract: Levenshtein distance of two sequences."""


def levenshtein(a, b) -> int:
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


# pylev exposes several aliases of the same function
classic_levenshtein = recursive_levenshtein = wf_levenshtein = wfi_levenshtein = damerau_levenshtein = levenshtein
