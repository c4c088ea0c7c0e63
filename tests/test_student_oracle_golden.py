"""CPU: the student-trainer oracle (oracle/student_oracle.py) against the goldens generated from the live reference
(oracle/make_golden_student.py -> tests/golden/student.pt)."""
import os

import pytest
import torch

from oracle import student_oracle as S
from oracle.make_golden import summarize

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'student.pt')


@pytest.mark.parametrize('case', ['small', 'mid'])
def test_student_oracle_matches_reference_golden(case):
    rec = torch.load(GOLD)[case]
    cfg = S.student_cfg(**rec['spec']['cfg'])
    prm = S.make_student_params(cfg, seed=0)
    batch = S.make_student_batch(cfg, rec['spec']['B'], seed=1234)
    mon, grads, newp, idx = S.student_step(batch, prm, cfg, rec['masked_event_index'], lr=rec['lr'])
    for k in ('loss_teacher', 'loss_quantization', 'loss_reconstruction', 'loss_encdec'):
        assert abs(float(mon[k]) - rec['means'][k]) <= 1e-5 * abs(rec['means'][k]), k
    assert torch.equal(idx, rec['indices'])                      # code indices: bit-exact
    assert set(prm) == set(rec['delta'])                         # parameter names = the reference's state_dict keys
    for k in prm:
        d, r = summarize(newp[k] - prm[k], k), rec['delta'][k]
        assert abs(d['norm'] - r['norm']) <= 2e-3 * max(r['norm'], 1e-12), (k, d['norm'], r['norm'])
        assert abs(d['proj'] - r['proj']) <= 2e-3 * max(r['norm'], 1e-12) * max(1.0, d['numel'] ** 0.5), k
