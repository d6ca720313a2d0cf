"""Loader for maneuver automata exported by the AROC toolbox (restates reference
maneuverAutomaton/loadAROCautomaton.py:14-145; the exported ``automaton.zip`` itself is a large blob that is not
shipped with the reference sources).

Layout read:  <zip>/automaton.xml            root children: text = file name of one primitive
              <zip>/primitives/<name>.xml    <successors><id>..</id></successors>  (1-based)
                                             <trajectory><states><state><time|x|y|velocity|orientation/>..</states>
                                                         <inputs><input><acceleration|steer/>..</inputs></trajectory>
                                             <occupancySet><occupancy><time><intervalStart/><intervalEnd/></time>
                                                           <set><point><dim dimension="1|2"/>..</point>..</set>
The archive is unpacked next to it once and re-unpacked when its modification time changes (reference :17-41).
Occupancies carry 'starttime'/'endtime' (time *intervals*), which makes the planners test two consecutive free spaces
per occupancy (src/lowLevelPlannerManeuverAutomaton.py:25-30).
"""
import os
import pickle
import shutil
import time as timer
import xml.etree.ElementTree as ET

import numpy as np

from ..geometry import Polygon
from .ManeuverAutomaton import ManeuverAutomaton, MotionPrimitive

_STATE_ROW = {'x': 0, 'y': 1, 'velocity': 2, 'orientation': 3}
_INPUT_ROW = {'acceleration': 0, 'steer': 1}


def _unpack_if_changed(zip_path, target):
    stamp_file = os.path.join(target, 'lastVersion.obj')
    stamp = timer.ctime(os.path.getmtime(zip_path))
    if os.path.isdir(target) and os.path.isfile(stamp_file):
        with open(stamp_file, 'rb') as f:
            if pickle.load(f).get('time') == stamp:
                return
    if os.path.isdir(target):
        shutil.rmtree(target)
    shutil.unpack_archive(zip_path, target)
    with open(stamp_file, 'wb') as f:
        pickle.dump({'time': stamp}, f)


def _parse_primitive(path):
    root = ET.parse(path).getroot()
    successors, x, u, t, occ = [], None, None, None, []
    for part in root:
        if part.tag == 'successors':
            successors = [int(e.text) - 1 for e in part]
        elif part.tag == 'trajectory':
            for block in part:
                if block.tag == 'states':
                    x = np.zeros((4, len(block)))
                    t = np.zeros(len(block))
                    for k, state in enumerate(block):
                        for e in state:
                            if e.tag == 'time':
                                t[k] = float(e.text)
                            elif e.tag in _STATE_ROW:
                                x[_STATE_ROW[e.tag], k] = float(e.text)
                elif block.tag == 'inputs':
                    u = np.zeros((2, len(block)))
                    for k, inp in enumerate(block):
                        for e in inp:
                            if e.tag in _INPUT_ROW:
                                u[_INPUT_ROW[e.tag], k] = float(e.text)
        elif part.tag == 'occupancySet':
            for o in part:
                t0 = t1 = None
                V = None
                for e in o:
                    if e.tag == 'time':
                        for b in e:
                            if b.tag == 'intervalStart':
                                t0 = float(b.text)
                            elif b.tag == 'intervalEnd':
                                t1 = float(b.text)
                    elif e.tag == 'set':
                        V = np.zeros((len(e), 2))
                        for k, point in enumerate(e):
                            for d in point:
                                V[k, int(d.attrib['dimension']) - 1] = float(d.text)
                occ.append({'space': Polygon(V), 'starttime': t0, 'endtime': t1})
    return MotionPrimitive(x, u, t[-1], successors, occ)


def loadAROCautomaton(zip_path=None):
    """load a maneuver automaton created with the AROC toolbox"""
    if zip_path is None:
        zip_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'AROC', 'automaton.zip')
    if not os.path.isfile(zip_path):
        raise FileNotFoundError("%s not found: the AROC export is not shipped; synthesise it with the reference's "
                                "maneuverAutomaton/AROC/createManeuverAutomaton.m or use load_maneuver_automaton()"
                                % zip_path)
    target = os.path.join(os.path.dirname(zip_path), 'automaton')
    _unpack_if_changed(zip_path, target)
    index = ET.parse(os.path.join(target, 'automaton.xml')).getroot()
    prims = [_parse_primitive(os.path.join(target, 'primitives', child.text)) for child in index]
    listed = [p.successors for p in prims]
    MA = ManeuverAutomaton(prims)
    MA.listed_successors = listed       # the export's own successor lists (the constructor derives them from velocities)
    return MA
