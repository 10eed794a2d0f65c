"""The closed-form stand-in 'UNet' used when the golden log_validation fixture was generated
(oracle/make_golden.py).  Kept separate so tests can import it without importing /root/reference.
TEST INFRASTRUCTURE ONLY."""
import torch


def stub_eps(lat, t):
    return torch.tanh(0.7 * lat + 0.1 * lat.roll(1, -1)) * (0.5 + 0.0004 * float(t))
