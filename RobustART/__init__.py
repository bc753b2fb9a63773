"""Drop-in import path for the B200-native replacement of RobustART's noise + eval hot path.

Only the hot-path surface is provided: `RobustART.noise.AddNoise` and (via `prototype.prototype.solver`)
the evaluation command lines.  Everything is backed by robustart_b200 (hand-written sm_100a CUDA)."""
