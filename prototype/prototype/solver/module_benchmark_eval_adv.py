"""Same CLI as benchmark_eval_adv; the reference's variant differs only in checkpoint-key munging
(robustart_b200.nets._strip_prefix handles every variant's prefixes)."""
from prototype.prototype.solver.benchmark_eval_adv import main  # noqa: F401

if __name__ == "__main__":
    main()
