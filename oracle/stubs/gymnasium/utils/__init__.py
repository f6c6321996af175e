from gym.utils import EzPickle, seeding
