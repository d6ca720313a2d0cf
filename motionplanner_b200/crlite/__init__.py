"""crlite -- a minimal reader of CommonRoad scenario files (2018b / 2020a XML) with the attribute names of
``commonroad-io`` that the collision-check path touches.

The reference depends on ``commonroad-io==2023.1`` (requirements.txt:1), which is not a dependency of this package.
Only what the hot path and its callers read is restated: lanelet bounds and adjacency
(auxiliary/polygonLaneletNetwork.py:5-129), obstacle shapes and ``occupancy_at_time``
(auxiliary/collisionChecker.py:25-33, src/maneuverAutomatonPlannerStandalone.py:165-180), the planning problem's
initial state and goal (src/maneuverAutomatonPlannerStandalone.py:98-149, src/highLevelPlanner.py:74-198).
Objects from the real ``commonroad`` package are accepted everywhere these are (duck typing on the same names).
"""
import xml.etree.ElementTree as ET

import numpy as np

from ..geometry import Point, Polygon as _GPolygon, point_in_rings

TWO_PI = 2.0 * np.pi


def make_valid_orientation(angle):
    """commonroad.common.util.make_valid_orientation: wrap to [-pi, pi]"""
    angle = angle % TWO_PI
    if np.pi <= angle <= TWO_PI:
        angle = angle - TWO_PI
    return angle


def rotate_translate(vertices, translation, angle):
    """commonroad.geometry.transform.rotate_translate: homogeneous [[c,-s,tx],[s,c,ty]] applied to every vertex"""
    c, s = np.cos(angle), np.sin(angle)
    v = np.asarray(vertices, dtype=np.float64)
    x = c * v[:, 0] + (-s) * v[:, 1] + translation[0]
    y = s * v[:, 0] + c * v[:, 1] + translation[1]
    return np.stack([x, y], axis=1)


class Interval:
    def __init__(self, start, end):
        self.start, self.end = start, end

    def contains(self, v):
        return self.start <= v <= self.end


class Shape:
    @property
    def shapely_object(self):
        return _GPolygon(self.vertices)


class Rectangle(Shape):
    def __init__(self, length, width, center=None, orientation=0.0):
        self.length, self.width = float(length), float(width)
        self.center = np.array([0.0, 0.0]) if center is None else np.asarray(center, dtype=np.float64)
        self.orientation = float(orientation)
        l2, w2 = 0.5 * self.length, 0.5 * self.width
        local = np.array([[-l2, -w2], [-l2, w2], [l2, w2], [l2, -w2], [-l2, -w2]])
        self.vertices = rotate_translate(local, self.center, self.orientation)

    def rotate_translate_local(self, translation, angle):
        return Rectangle(self.length, self.width, self.center + np.asarray(translation),
                         make_valid_orientation(self.orientation + angle))


class Polygon(Shape):
    def __init__(self, vertices):
        v = np.asarray(vertices, dtype=np.float64)
        if len(v) and not np.array_equal(v[0], v[-1]):
            v = np.concatenate([v, v[:1]])
        self.vertices = v

    @property
    def center(self):
        return self.vertices[:-1].mean(axis=0)

    def rotate_translate_local(self, translation, angle):
        c = self.center
        return Polygon(rotate_translate(self.vertices - c, c + np.asarray(translation), angle))

    def contains_point(self, p):
        return point_in_rings((float(p[0]), float(p[1])), [[tuple(q) for q in self.vertices]]) >= 0


class ShapeGroup(Shape):
    def __init__(self, shapes):
        self.shapes = list(shapes)

    def rotate_translate_local(self, translation, angle):
        return ShapeGroup([s.rotate_translate_local(translation, angle) for s in self.shapes])

    @property
    def shapely_object(self):
        raise NotImplementedError("a ShapeGroup is consumed piecewise (reference get_shapely_object unions them)")


class State:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def attributes(self):
        return list(self.__dict__.keys())


class Occupancy:
    def __init__(self, time_step, shape):
        self.time_step, self.shape = time_step, shape


class TrajectoryPrediction:
    def __init__(self, states, shape):
        self.trajectory_states = states
        self.shape = shape
        self.initial_time_step = states[0].time_step
        self.final_time_step = states[-1].time_step
        self._by_time = {s.time_step: s for s in states}

    def occupancy_at_time_step(self, t):
        s = self._by_time.get(t)
        if s is None:
            return None
        return Occupancy(t, self.shape.rotate_translate_local(s.position, s.orientation))


class SetBasedPrediction:
    def __init__(self, occupancies):
        self.occupancy_set = occupancies
        steps = [o.time_step if not isinstance(o.time_step, Interval) else o.time_step.start for o in occupancies]
        ends = [o.time_step if not isinstance(o.time_step, Interval) else o.time_step.end for o in occupancies]
        self.initial_time_step, self.final_time_step = min(steps), max(ends)

    def occupancy_at_time_step(self, t):
        for o in self.occupancy_set:
            ts = o.time_step
            if (isinstance(ts, Interval) and ts.contains(t)) or ts == t:
                return o
        return None


class Obstacle:
    def __init__(self, obstacle_id, role, obstacle_type, shape, initial_state, prediction=None):
        self.obstacle_id, self.obstacle_role, self.obstacle_type = obstacle_id, role, obstacle_type
        self.obstacle_shape, self.initial_state, self.prediction = shape, initial_state, prediction
        self._initial_occupancy = Occupancy(initial_state.time_step, shape.rotate_translate_local(
            initial_state.position, initial_state.orientation))

    def occupancy_at_time(self, t):
        """commonroad Obstacle.occupancy_at_time: static obstacles occupy every step; dynamic ones the initial shape
        at the initial step, the prediction inside its range, None otherwise"""
        if self.obstacle_role == 'static':
            return self._initial_occupancy
        if t == self.initial_state.time_step:
            return self._initial_occupancy
        p = self.prediction
        if p is not None and p.initial_time_step <= t <= p.final_time_step:
            return p.occupancy_at_time_step(t)
        return None


class Lanelet:
    def __init__(self, lanelet_id, left, right):
        self.lanelet_id = lanelet_id
        self.left_vertices, self.right_vertices = left, right
        self.center_vertices = 0.5 * (left + right)
        self.predecessor, self.successor = [], []
        self.adj_left = self.adj_right = None
        self.adj_left_same_direction = self.adj_right_same_direction = None
        self.polygon = Polygon(np.concatenate([right, np.flip(left, 0)]))
        # commonroad Lanelet.distance: arc length of the centre line up to every centre vertex
        self.distance = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(self.center_vertices, axis=0), axis=1))])

    def orientation_by_position(self, position):
        """direction of the centre-line segment closest to `position` (commonroad Lanelet.orientation_by_position)"""
        c = self.center_vertices
        a, d = c[:-1], c[1:] - c[:-1]
        n2 = np.maximum((d * d).sum(axis=1), 1e-300)
        t = np.clip(((np.asarray(position) - a) * d).sum(axis=1) / n2, 0.0, 1.0)
        k = int(np.argmin(np.linalg.norm(a + t[:, None] * d - np.asarray(position), axis=1)))
        return float(np.arctan2(d[k, 1], d[k, 0]))


class LaneletNetwork:
    def __init__(self, lanelets):
        self.lanelets = lanelets
        self._by_id = {l.lanelet_id: l for l in lanelets}

    def find_lanelet_by_id(self, lanelet_id):
        return self._by_id.get(lanelet_id)

    def find_lanelet_by_position(self, points):
        return [[l.lanelet_id for l in self.lanelets if l.polygon.contains_point(p)] for p in points]

    def find_most_likely_lanelet_by_state(self, state_list):
        """per state the lanelet that contains its position and whose direction there is closest to the state's
        orientation (commonroad LaneletNetwork.find_most_likely_lanelet_by_state); states off the road give nothing"""
        out = []
        for st in state_list:
            ids = self.find_lanelet_by_position([st.position])[0]
            if not ids:
                continue
            diff = [abs(make_valid_orientation(self._by_id[i].orientation_by_position(st.position) - st.orientation))
                    for i in ids]
            out.append(ids[int(np.argmin(diff))])
        return out


class Scenario:
    def __init__(self, dt, scenario_id, lanelet_network, static_obstacles, dynamic_obstacles, tags):
        self.dt, self.scenario_id, self.lanelet_network = dt, scenario_id, lanelet_network
        self.static_obstacles, self.dynamic_obstacles, self.tags = static_obstacles, dynamic_obstacles, tags

    @property
    def obstacles(self):
        return list(self.static_obstacles) + list(self.dynamic_obstacles)

    # the part of commonroad's Scenario editing interface auxiliary/prediction.py uses (:41-42, :166-172)
    def remove_obstacle(self, obstacle):
        for lst in (self.static_obstacles, self.dynamic_obstacles):
            if obstacle in lst:
                lst.remove(obstacle)

    def add_objects(self, obj):
        for o in (obj if isinstance(obj, (list, tuple)) else [obj]):
            (self.static_obstacles if o.obstacle_role == 'static' else self.dynamic_obstacles).append(o)

    def generate_object_id(self):
        ids = [o.obstacle_id for o in self.obstacles] + [l.lanelet_id for l in self.lanelet_network.lanelets]
        self._next_id = max([getattr(self, '_next_id', 0)] + ids) + 1
        return self._next_id


class GoalRegion:
    def __init__(self, state_list, lanelets_of_goal_position):
        self.state_list, self.lanelets_of_goal_position = state_list, lanelets_of_goal_position


class PlanningProblem:
    def __init__(self, planning_problem_id, initial_state, goal):
        self.planning_problem_id, self.initial_state, self.goal = planning_problem_id, initial_state, goal


class PlanningProblemSet:
    def __init__(self, problems):
        self.planning_problem_dict = {p.planning_problem_id: p for p in problems}


# ---------------------------------------------------------------------------------------------------------------------
def _f(e, path):
    n = e.find(path)
    return None if n is None else float(n.text)


def _points(e):
    return np.array([[float(p.find('x').text), float(p.find('y').text)] for p in e.findall('point')])


def _value(e):
    """exact value or Interval of a state element"""
    if e is None:
        return None
    ex = e.find('exact')
    if ex is not None:
        return float(ex.text)
    return Interval(float(e.find('intervalStart').text), float(e.find('intervalEnd').text))


def _time(e):
    v = _value(e)
    if isinstance(v, Interval):
        return Interval(int(v.start), int(v.end))
    return int(v)


def _shape(e):
    """<shape> / <position> content -> Shape (several children -> ShapeGroup)"""
    shapes = []
    for ch in e:
        if ch.tag == 'rectangle':
            center = ch.find('center')
            c = None if center is None else [float(center.find('x').text), float(center.find('y').text)]
            shapes.append(Rectangle(_f(ch, 'length'), _f(ch, 'width'), c, _f(ch, 'orientation') or 0.0))
        elif ch.tag == 'polygon':
            shapes.append(Polygon(_points(ch)))
        elif ch.tag == 'circle':
            raise NotImplementedError("circle shapes are not supported by crlite")
    if not shapes:
        return None
    return shapes[0] if len(shapes) == 1 else ShapeGroup(shapes)


def _state(e, goal=False, lanelets=None):
    kw = {}
    for ch in e:
        if ch.tag == 'position':
            pt = ch.find('point')
            if pt is not None:
                kw['position'] = np.array([float(pt.find('x').text), float(pt.find('y').text)])
            else:
                refs = [int(l.attrib['ref']) for l in ch.findall('lanelet')]
                if refs:
                    polys = [lanelets[r].polygon for r in refs]
                    kw['position'] = polys[0] if len(polys) == 1 else ShapeGroup(polys)
                    kw['_lanelet_refs'] = refs
                else:
                    kw['position'] = _shape(ch)
        elif ch.tag == 'time':
            kw['time_step'] = _time(ch)
        elif ch.tag in ('orientation', 'velocity', 'acceleration', 'yawRate', 'slipAngle'):
            name = {'yawRate': 'yaw_rate', 'slipAngle': 'slip_angle'}.get(ch.tag, ch.tag)
            kw[name] = _value(ch)
    if goal and not isinstance(kw.get('time_step'), Interval):
        t = kw.get('time_step', 0)
        kw['time_step'] = Interval(t, t)
    return State(**kw)


def _obstacle(e, role_default=None):
    role = e.findtext('role') or role_default or ('static' if e.tag == 'staticObstacle' else 'dynamic')
    shape = _shape(e.find('shape'))
    init = _state(e.find('initialState'))
    if not hasattr(init, 'orientation'):
        init.orientation = 0.0
    pred = None
    tr = e.find('trajectory')
    if tr is not None:
        states = [_state(s) for s in tr.findall('state')]
        if states:
            pred = TrajectoryPrediction(states, shape)
    occ = e.find('occupancySet')
    if occ is not None:
        pred = SetBasedPrediction([Occupancy(_time(o.find('time')), _shape(o.find('shape'))) for o in occ.findall('occupancy')])
    return Obstacle(int(e.attrib['id']), role, e.findtext('type'), shape, init, pred)


class CommonRoadFileReader:
    def __init__(self, filename):
        self.filename = filename

    def open(self, lanelet_assignment=False):
        root = ET.parse(self.filename).getroot()
        dt = float(root.attrib['timeStepSize'])
        lanelets = []
        for e in root.findall('lanelet'):
            l = Lanelet(int(e.attrib['id']), _points(e.find('leftBound')), _points(e.find('rightBound')))
            l.predecessor = [int(p.attrib['ref']) for p in e.findall('predecessor')]
            l.successor = [int(p.attrib['ref']) for p in e.findall('successor')]
            a = e.find('adjacentLeft')
            if a is not None:
                l.adj_left, l.adj_left_same_direction = int(a.attrib['ref']), a.attrib['drivingDir'] == 'same'
            a = e.find('adjacentRight')
            if a is not None:
                l.adj_right, l.adj_right_same_direction = int(a.attrib['ref']), a.attrib['drivingDir'] == 'same'
            lanelets.append(l)
        net = LaneletNetwork(lanelets)
        static, dynamic = [], []
        for e in root:
            if e.tag in ('obstacle', 'dynamicObstacle', 'staticObstacle'):
                o = _obstacle(e)
                (static if o.obstacle_role == 'static' else dynamic).append(o)
        tags = (root.attrib.get('tags') or '').split()
        st = root.find('scenarioTags')
        if st is not None:
            tags = [t.tag for t in st]
        scenario = Scenario(dt, root.attrib.get('benchmarkID', ''), net, static, dynamic, tags)
        problems = []
        for e in root.findall('planningProblem'):
            init = _state(e.find('initialState'))
            goals = [_state(g, goal=True, lanelets=net._by_id) for g in e.findall('goalState')]
            refs = {i: g.__dict__.pop('_lanelet_refs') for i, g in enumerate(goals) if '_lanelet_refs' in g.__dict__}
            problems.append(PlanningProblem(int(e.attrib['id']), init, GoalRegion(goals, refs or None)))
        return scenario, PlanningProblemSet(problems)
