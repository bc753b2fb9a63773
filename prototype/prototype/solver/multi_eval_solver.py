"""python -m prototype.prototype.solver.multi_eval_solver --config C --evaluate [--ckpt-filePath DIR]
(reference: multi_eval_solver.py main loop)

Clean-accuracy evaluation of every model named in config.eval_list (or of config.model), one after the other, on the
same sharded validation set; per-model metrics under `<save_path>/<name>/results/`.  Load failures are logged to
status.txt and skipped like the reference does."""
import argparse
import os
import traceback

from robustart_b200 import solver as S


def main(argv=None):
    parser = argparse.ArgumentParser(description="Classification Solver")
    parser.add_argument("--config", required=True, type=str)
    parser.add_argument("--evaluate", action="store_true")
    parser.add_argument("--ckpt-filePath", default="/mnt/lustre/share/robust/ckpt_baseline")
    args = parser.parse_args(argv)
    if not args.evaluate:
        raise SystemExit("training is outside the B200 hot path; run with --evaluate")
    config = S.parse_config(args.config)
    d = S.dist_init()
    results = {}
    status = open("status.txt", "w") if d.rank == 0 else None
    names = list(config["eval_list"]) if "eval_list" in config else [None]
    for name in names:
        try:
            if name is None:
                cfg, label = config.model, config.model["type"]
                pre = config.get("saver", {}).get("pretrain", {})
                ckpt = pre.get("path") if pre else None
            else:
                cfg, label, ckpt = S.model_name_dict[name], name, os.path.join(args.ckpt_filePath, name + ".pth.tar")
            sol = S.EvalSolver(config, prefix=label if name else "", dist_info=d)
            model = S.build_b200_model(cfg, ckpt, sol.device)
        except Exception:
            print("Error when load %s" % name)
            print(traceback.format_exc())
            if status:
                status.write("Error when load %s, skip it.\n%s" % (name, traceback.format_exc()))
            continue
        results[label] = sol.evaluate(model)
        if status:
            status.write("%s done\n" % label)
    if status:
        status.close()
    return results


if __name__ == "__main__":
    main()
