"""Extract the pinned values of the reference's own block-GMG tests into tests/golden/reference_screen_output.json.

Reads the deal.II 9.6 `screen-output` goldens under /root/reference/tests; run it in the build
container only -- /root/reference does not exist on the GPU box, the JSON it writes is what the tests read.
    python tools/extract_goldens.py
"""
import json
import os
import re

REF = "/root/reference/tests"
TESTS = ["nsinker_gmg_no_avg", "nsinker_gmg_gc", "nsinker_gmg_gc_no_avg", "nsinker_gmg_idr", "sol_cx_2_gmg", "sol_cx_4_gmg",
         "sol_cx_4_gmg_no_avg", "no_cheap_gmg", "performance/performance_nsinker_gmg_idr", "spherical_nsinker_gmg_gc", "hollow_sphere_gmg",
         "hollow_sphere_gmg_project_q1", "prescribed_dilation_gmg", "shell_3d_gmg", "gmg_no_normal_flux_only", "muparser_temperature_example", "initial_global_refinement"]


def num(s):
    return int(s.replace(",", ""))


def parse(path):
    txt = open(path).read()
    out = {"file": path.replace("/root/reference/", "")}
    out["active_cells"] = [num(m) for m in re.findall(r"Number of active cells: ([\d,]+)", txt)]
    out["levels"] = [int(m) for m in re.findall(r"\(on (\d+) levels\)", txt)]
    out["dofs_u_p"] = [[num(a), num(b)] for a, b in re.findall(r"degrees of freedom: [\d,]+ \(([\d,]+)\+([\d,]+)", txt)]
    out["coarse_sizes_A_S"] = [[int(a), int(b)] for a, b in re.findall(r"GMG coarse size A: (\d+), coarse size S: (\d+)", txt)]
    out["viscosity_range"] = [[float(a), float(b)] for a, b in re.findall(r"Viscosity range: (\S+) - (\S+)", txt)]
    out["stokes_iterations_cheap_expensive"] = [[int(a), int(b)] for a, b in
                                                 re.findall(r"(?:Stokes solver \([^:]*\): |Solving Stokes system \(GMG\)\.\.\. )(\d+)\+(\d+) iterations", txt)]
    out["schur_preconditioner_applications"] = [int(a) for a in re.findall(r"Schur complement preconditioner: (\d+)\+", txt)]
    out["A_preconditioner_applications"] = [int(a) for a in re.findall(r"A block preconditioner: (\d+)\+", txt)]
    m = re.findall(r"Errors u_L1, p_L1, u_L2, p_L2(?: topo_L2)?: (\S+), (\S+), (\S+), ([^\s,]+)", txt)
    out["errors_uL1_pL1_uL2_pL2"] = [list(x) for x in m]      # kept as the printed strings
    out["errors_uL2_pL2"] = [list(x) for x in re.findall(r"Errors u_L2, p_L2:\s+(\S+), (\S+)", txt)]
    return out


def main():
    res = {}
    for t in TESTS:
        res[os.path.basename(t)] = parse(os.path.join(REF, t, "screen-output"))
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_screen_output.json")
    with open(dst, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote", dst)


if __name__ == "__main__":
    main()
