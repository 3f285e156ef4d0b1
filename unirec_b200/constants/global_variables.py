# reference: unirec/constants/global_variables.py:4-6
EPS = 1e-8
VALID_TRIGGER_P = 0.1   # probability with which the BPR label-layout check runs (reco_abc.py:239-246)
