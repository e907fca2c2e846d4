# -*- coding: utf-8 -*-
from .common import *
from .cwt_utils import *
from .stft_utils import *
