"""Test infrastructure: CPU oracle of the ContraD hot path (see contrad_oracle.py header).
Never imported by the product package `contrad_b200`."""
