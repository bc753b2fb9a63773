"""python -m prototype.prototype.solver.benchmark_eval_adv --config C --src_name S --src_path P --tgt_name T
       --tgt_path Q --attack A --eps E          (reference: benchmark_eval_adv.py:261-299)

--eps is a Python expression such as 4/255, evaluated like the reference does (:295-299)."""
import argparse

from robustart_b200 import solver as S


def main(argv=None):
    parser = argparse.ArgumentParser(description="Classification Solver")
    parser.add_argument("--config", required=True, type=str)
    parser.add_argument("--src_name", required=True, type=str)
    parser.add_argument("--src_path", required=True, type=str)
    parser.add_argument("--tgt_name", required=True, type=str)
    parser.add_argument("--tgt_path", required=True, type=str)
    parser.add_argument("--attack", required=True, type=str)
    parser.add_argument("--eps", required=True, type=str)
    args = parser.parse_args(argv)
    config = S.parse_config(args.config)
    config.model_src = S.model_name_dict[args.src_name]
    config.model_tgt = S.model_name_dict[args.tgt_name]
    eps = eval(args.eps)  # noqa: S307 -- mirrors the reference CLI contract
    prefix = args.tgt_name + "_" + args.attack + "_" + str(eps)[0:min(5, len(str(eps)))]
    d = S.dist_init()
    sol = S.EvalSolver(config, prefix=prefix, dist_info=d)
    model_src = S.build_source_model(config.model_src, args.src_path, sol.device)
    model_tgt = S.build_b200_model(config.model_tgt, args.tgt_path, sol.device)
    return sol.evaluate_adv(model_src, model_tgt, attack=args.attack, eps=eps)


if __name__ == "__main__":
    main()
