"""vehicle parameter dictionary -- same keys and values as reference vehicle/vehicleParameter.py:3-15"""


def vehicleParameter():
    """parameter for the car"""
    wheelbase = 2.3
    return {
        'length': 4.3,              # length of the car
        'width': 1.7,               # width of the car
        'wheelbase': wheelbase,     # length of the wheelbase
        'b': wheelbase / 2,         # distance from car center to rear axis
        'a_max': 9,                 # maximum acceleration
        's_max': 1,                 # maximum steering angle
        'svel_max': 0.4,            # maximum steering angle velocity
    }
