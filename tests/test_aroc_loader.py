"""AROC automaton loader (reference maneuverAutomaton/loadAROCautomaton.py:14-145) on a synthetic export with the
same XML layout -- the real automaton.zip is a missing blob."""
import os
import shutil

import numpy as np

from motionplanner_b200.maneuverAutomaton.loadAROCautomaton import loadAROCautomaton


def _write_export(tmp, nprim=3):
    root = os.path.join(tmp, 'export')
    os.makedirs(os.path.join(root, 'primitives'))
    with open(os.path.join(root, 'automaton.xml'), 'w') as f:
        f.write('<automaton>' + ''.join('<primitive>p%d.xml</primitive>' % i for i in range(nprim)) + '</automaton>')
    for i in range(nprim):
        v0 = 0.4 * i
        states = ''.join('<state><time>%.1f</time><x>%.3f</x><y>0.0</y><velocity>%.1f</velocity><orientation>0.0'
                         '</orientation></state>' % (0.1 * k, v0 * 0.1 * k, v0 if k < 10 else 0.4 * ((i + 1) % nprim))
                         for k in range(11))
        inputs = ''.join('<input><acceleration>0.5</acceleration><steer>0.0</steer></input>' for _ in range(10))
        occ = ''
        for k in range(10):
            pts = [(v0 * 0.1 * k - 1.0, -0.9), (v0 * 0.1 * k + 3.4, -0.9), (v0 * 0.1 * k + 3.6, 0.0),
                   (v0 * 0.1 * k + 3.4, 0.9), (v0 * 0.1 * k - 1.0, 0.9)]
            occ += '<occupancy><time><intervalStart>%.1f</intervalStart><intervalEnd>%.1f</intervalEnd></time><set>' % (0.1 * k, 0.1 * k + 0.1)
            occ += ''.join('<point><dim dimension="1">%r</dim><dim dimension="2">%r</dim></point>' % p for p in pts)
            occ += '</set></occupancy>'
        with open(os.path.join(root, 'primitives', 'p%d.xml' % i), 'w') as f:
            f.write('<primitive><successors><id>%d</id><id>%d</id></successors><trajectory><states>%s</states><inputs>%s'
                    '</inputs></trajectory><occupancySet>%s</occupancySet></primitive>' % (1, (i + 1) % nprim + 1, states, inputs, occ))
    os.makedirs(os.path.join(tmp, 'AROC'))
    return shutil.make_archive(os.path.join(tmp, 'AROC', 'automaton'), 'zip', root)


def test_load_synthetic_aroc_export(tmp_path):
    z = _write_export(str(tmp_path))
    MA = loadAROCautomaton(z)
    assert len(MA.primitives) == 3 and not MA.is_timepoint()
    p = MA.primitives[1]
    assert p.x.shape == (4, 11) and p.u.shape == (2, 10) and abs(p.tFinal - 1.0) < 1e-12
    assert MA.listed_successors[1] == [0, 2]                     # 1-based ids -> 0-based
    assert len(p.occ) == 10 and p.occ[3]['starttime'] == 0.3 and len(p.occ[0]['space'].exterior.coords) == 6
    lay = MA.device_layout(0.1)
    # int(starttime/0.1): 0.3 -> 2, 0.6 -> 5, 0.7 -> 6 (float truncation, quirk Q2) => duplicated / skipped indices
    want = sorted(int(float('%.1f' % (0.1 * k)) / 0.1) for k in range(10))   # values as parsed from decimal text
    assert sorted(lay['tidx'][0].tolist()) == want and want.count(2) == 2 and 3 not in want
    assert lay['verts'].shape[2] == 5 and (lay['nv'] == 5).all()
    # unpack cache: second load reuses the extracted directory, a touched zip re-extracts
    marker = os.path.join(os.path.dirname(z), 'automaton', 'marker')
    open(marker, 'w').close()
    loadAROCautomaton(z)
    assert os.path.exists(marker)
    os.utime(z, (1, 1))
    loadAROCautomaton(z)
    assert not os.path.exists(marker)
