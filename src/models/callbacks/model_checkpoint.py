from peclr_b200.lightning import ModelCheckpoint as UpdatedModelCheckpoint  # noqa: F401
