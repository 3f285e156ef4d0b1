"""Generate tests/golden_eval/onepos_metrics.npz from the UNMODIFIED reference evaluator (TEST INFRASTRUCTURE ONLY).

    python oracle/make_eval_golden.py        # in the build container, where /root/reference exists

Random one-positive score matrices [n, 1+K] (positive in column 0, continuous scores: no ties) go through
`OnePositiveEvaluator.evaluate_with_scores` (unirec/facility/evaluation/onepos.py:104-175); the per-sample metric vectors it returns
are averaged like `Evaluator.merge_scores` does.  Only shim: `np.Inf` (removed in numpy 2), which the reference uses as "no cutoff".
"""
import os
import sys

import numpy as np

REF = os.environ.get('UNIREC_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden_eval')
METRICS = "['hit@1;5;10', 'ndcg@5;10', 'mrr@5', 'mrr', 'ndcg', 'group_auc']"


def main():
    sys.path.insert(0, REF)
    if not hasattr(np, 'Inf'):
        np.Inf = np.inf
    from unirec.facility.evaluation.onepos import OnePositiveEvaluator
    rng = np.random.RandomState(7)
    out = {'metrics_str': np.frombuffer(METRICS.encode(), dtype=np.uint8)}
    for tag, (n, N) in {'a': (257, 12), 'b': (64, 101)}.items():
        scores = rng.randn(n, N).astype(np.float64)
        ev = OnePositiveEvaluator(METRICS)
        res = ev.evaluate_with_scores(scores.copy())
        out['scores_' + tag] = scores
        for k, v in res.items():
            out['metric_%s/%s' % (tag, k)] = np.asarray(np.mean(v), dtype=np.float64)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, 'onepos_metrics.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, sorted(k for k in out if k.startswith('metric_a')))


if __name__ == '__main__':
    main()
