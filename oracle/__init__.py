"""TEST INFRASTRUCTURE ONLY.

CPU restatement (plain PyTorch fp32) of the reference's segmentation step, used as the parity
checker by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  Nothing under ``sim2real_lane_segment_b200/`` may import it.

Parity pin: the reference (timurlenk07/sim2real_lane_segment) ships no tests or golden vectors
(SURVEY.md §4), so the pin is the reference's own model code imported in the build container:
``oracle/make_golden.py`` runs ``/root/reference/rightLaneNetwork/models/FCDenseNet`` on seeded inputs
and stores the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures (and, when ``/root/reference`` is present, against the live reference modules).
The training modules themselves need ``pytorch_lightning==1.2.1`` (not installable here), so
``SimpleTrainModule.training_step`` / ``evaluate_batch`` and the four PL metric functions are restated
from the published PL-1.2.1 semantics: for those the pin is "restated, reference not runnable".
"""
