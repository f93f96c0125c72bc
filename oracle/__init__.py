"""Test infrastructure only — see qmatmul_oracle.py.  Never imported by chatglm_q_b200/."""
