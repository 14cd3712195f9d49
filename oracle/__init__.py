"""TEST INFRASTRUCTURE: CPU restatement of the reference hot path (see t2s_oracle.py).  Never imported by the product package."""
