"""Batch harness over a directory of CommonRoad scenarios for the accelerated path -- writes the same CSV the
reference's evaluation/runAllScenarios.py:175,181 writes (``file, comp_time/final_time | 'failed' | error, 'collision' |
'no collision', tags, final_time``), so the reference's evaluation/evaluationScript.py aggregates it unchanged.

Only the planner that needs no decision module is available here (PLANNER = 'AutomatonStandalone', reference :88-90);
'Automaton' needs ``highLevelPlanner`` (out of scope) -- pass its outputs to
``lowLevelPlannerManeuverAutomaton`` directly instead.

    python -m motionplanner_b200.evaluation.runAllScenarios <scenario dir> [results dir] [max popped nodes]
"""
import os
import sys
import time

import numpy as np

from ..auxiliary.collisionChecker import collisionChecker
from ..crlite import CommonRoadFileReader
from ..maneuverAutomaton.createManeuverAutomaton import load_maneuver_automaton
from ..src.maneuverAutomatonPlannerStandalone import maneuverAutomatonPlannerStandalone
from ..vehicle.vehicleParameter import vehicleParameter

PLANNER = 'AutomatonStandalone'


def solve_scenario(path, MA, max_nodes=20000, backend=None, validator=None):
    """one row of the result table (reference solve_scenario :37-108)"""
    scenario, planning_problem = CommonRoadFileReader(path).open(lanelet_assignment=True)
    return solve_loaded(os.path.basename(path), scenario, planning_problem, MA, max_nodes, backend, validator)[0]


def solve_loaded(file, scenario, planning_problem, MA, max_nodes=20000, backend=None, validator=None):
    """solve_scenario for an already parsed scenario; returns (row, x, u, param)"""
    x = u = None
    pp = list(planning_problem.planning_problem_dict.values())[0]
    steps = 0
    for g in pp.goal.state_list:
        steps = np.maximum(g.time_step.end - pp.initial_state.time_step, steps)
    final_time = steps * scenario.dt
    tags = ';'.join(str(t) for t in scenario.tags)
    param = vehicleParameter()
    collision = 'no collision'
    try:
        start = time.time()
        x, u, _ = maneuverAutomatonPlannerStandalone(scenario, planning_problem, param, MA, backend=backend,
                                                     max_nodes=max_nodes)
        comp_time = time.time() - start
        if x is None:
            comp_time = 'failed'
        else:
            comp_time = comp_time / final_time
            ok = collisionChecker(scenario, x, param) if validator is None else validator(scenario, x, param)
            if not ok:
                collision = 'collision'
    except Exception as e:          # same convention as the reference: the message goes into the table
        comp_time = str(e).replace("\n", "").replace(",", "")
    return [file, comp_time, collision, tags, final_time], x, u, param


def run_all(scenario_dir, results_dir='results', max_nodes=20000, files=None, backend=None, validator=None):
    MA = load_maneuver_automaton()
    files = files or sorted(f for f in os.listdir(scenario_dir) if f.endswith('.xml'))
    data = [solve_scenario(os.path.join(scenario_dir, f), MA, max_nodes, backend, validator) for f in files]
    os.makedirs(results_dir, exist_ok=True)
    out = os.path.join(results_dir, 'computation_time_' + PLANNER + '.csv')
    np.savetxt(out, np.asarray(data, dtype=object), delimiter=",", fmt="%s")
    solved = sum(1 for d in data if not isinstance(d[1], str))
    print('success rate: ' + str(solved / max(len(files), 1) * 100) + "%")
    return out, data


if __name__ == "__main__":
    run_all(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'results',
            int(sys.argv[3]) if len(sys.argv) > 3 else 20000)
