"""Import-path compatibility with the reference repository: ``src.experiments.peclr_training``,
``src.models.*`` resolve to the B200-native implementation in ``peclr_b200``."""
