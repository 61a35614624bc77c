"""Module paths mirroring the reference's `model-module` strings (conf/*.yaml:3) so that
`dynamic_import("fcl_taco2_b200." + model_module)` resolves to the B200 classes."""
