import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'xd-tts_b200')
import numpy as np
from xdtts_b200 import griffin_lim as gl
from oracle import gl_oracle as o
basis=o.create_mel_filter_bank(22050.0,1024,80,0.0,8000.0)
for rf in (19, 0):
    voc=gl.GriffinLim.new(basis,768,1.7,60,0.99,run_frames=rf)
    mels=[o.synth_mel(1234+i,80,1000) for i in range(32)]
    try:
        ys=voc.infer_batch(mels); print('rf',rf,'ok',np.abs(ys[0]).max())
    except Exception as e:
        print('rf',rf,'FAIL',e)
