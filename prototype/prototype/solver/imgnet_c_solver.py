"""python -m prototype.prototype.solver.imgnet_c_solver --config C --evaluate [--ckpt-filePath DIR] [--save-detail]
(reference: imgnet_c_solver.py:477-541)

Evaluates one model (config.model) or every name in config.eval_list (checkpoints `<ckpt-filePath>/<name>.pth.tar`)
on the full ImageNet-C grid: 19 corruptions x 5 severities.  The corruptions are generated on the GPU from the clean
shard (RobustART.noise kernels) instead of being read from 95 pre-corrupted dataset copies; results land in
`<save_path>/<name>/results/{group}-{type}-{sev}-metric` and `robust.json` like the reference's merge_eval_res.
A model that fails to load is logged to status.txt and skipped, as in the reference's main loop."""
import argparse
import os
import shutil
import traceback

from robustart_b200 import solver as S


def main(argv=None):
    parser = argparse.ArgumentParser(description="Classification Solver")
    parser.add_argument("--config", required=True, type=str)
    parser.add_argument("--evaluate", action="store_true")
    parser.add_argument("--save-detail", action="store_true")
    parser.add_argument("--ckpt-filePath", default="/mnt/lustre/share/robust/ckpt_baseline")
    args = parser.parse_args(argv)
    if not args.evaluate:
        raise SystemExit("training is outside the B200 hot path; run with --evaluate")
    config = S.parse_config(args.config)
    d = S.dist_init()
    results = {}
    status = open("status.txt", "w") if d.rank == 0 else None

    def note(msg):
        if status:
            status.write(msg)

    if "eval_list" in config:
        for name in config["eval_list"]:
            try:
                cfg = S.model_name_dict[name]
                sol = S.EvalSolver(config, prefix=name, dist_info=d)
                model = S.build_b200_model(cfg, os.path.join(args.ckpt_filePath, name + ".pth.tar"), sol.device)
            except Exception:
                print("Error when load " + name)
                print(traceback.format_exc())
                note("Error when load %s, skip it.\n%s" % (name, traceback.format_exc()))
                continue
            results[name] = sol.evaluate_imagenet_c(model)
            note("%s done\n" % name)
            if not args.save_detail and d.rank == 0:          # keep robust.json, drop the per-cell files
                keep = os.path.join(sol.result_path, "robust.json")
                for f in os.listdir(sol.result_path):
                    if os.path.join(sol.result_path, f) != keep:
                        os.remove(os.path.join(sol.result_path, f))
    else:
        sol = S.EvalSolver(config, prefix="", dist_info=d)
        pre = config.get("saver", {}).get("pretrain", {})
        model = S.build_b200_model(config.model, pre.get("path") if pre else None, sol.device)
        results[config.model["type"]] = sol.evaluate_imagenet_c(model)
        note("%s done\n" % config.model["type"])
    if status:
        status.close()
    return results


if __name__ == "__main__":
    main()
