import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
	sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
	# a GPU test that hangs inside a CUDA call cannot be interrupted by a signal: let the watchdog thread end the run
	if config.pluginmanager.hasplugin("timeout"):
		config.option.timeout_method = "thread"
		if not getattr(config.option, "timeout", None):
			config.option.timeout = 600


@pytest.fixture(scope="session")
def golden():
	import numpy as np

	def load(name):
		return np.load(GOLDEN / f"{name}.npz")
	return load
