"""python -m prototype.prototype.solver.cls_solver --config C --evaluate   (reference: cls_solver.py:460-481)

Only evaluation is part of the hot path; without --evaluate the reference trains, which this framework
does not do.  YAML: model.{type,kwargs}, data.{batch_size,input_size,read_from,test.*}, saver.pretrain.path,
optional data.test.imagenet_c.{noise,severity} (or --corruption/--severity) for ImageNet-C evaluation with the
corruption generated on the GPU instead of read from a pre-corrupted dataset."""
import argparse

from robustart_b200 import solver as S


def main(argv=None):
    parser = argparse.ArgumentParser(description="Classification Solver")
    parser.add_argument("--config", required=True, type=str)
    parser.add_argument("--evaluate", action="store_true")
    parser.add_argument("--corruption", default=None, type=str)
    parser.add_argument("--severity", default=1, type=int)
    args = parser.parse_args(argv)
    if not args.evaluate:
        raise SystemExit("training is outside the B200 hot path; run with --evaluate")
    config = S.parse_config(args.config)
    d = S.dist_init()
    sol = S.EvalSolver(config, prefix="", dist_info=d)
    pre = config.get("saver", {}).get("pretrain", {})
    model = S.build_b200_model(config.model, pre.get("path") if pre else None, sol.device)
    c = config.data.get("test", {}).get("imagenet_c", {}) or {}
    corruption = args.corruption or c.get("noise")
    severity = args.severity if args.corruption else int(c.get("severity", args.severity))
    return sol.evaluate(model, corruption, severity)


if __name__ == "__main__":
    main()
