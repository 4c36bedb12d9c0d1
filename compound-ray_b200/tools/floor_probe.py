import sys, time, os
sys.path.insert(0,'/root/repo/compound-ray_b200'); sys.path.insert(0,'/root/repo')
import eye_renderer as er, numpy as np
sys.path.insert(0,'/root/repo/benchmarks')
import speed_test
data=speed_test.fixtures()
lib=er.load_library(device=0); lib.setVerbosity(False)
lib.loadGlTFscene(os.path.join(data,'data','natural-standin-sky.gltf').encode())
er.gotoFirstCompoundEye(lib)
er.setOmmatidiaFromOmmatidiumList(lib, er.readEyeFile(os.path.join(data,'data','eyes','1000-equidistant.eye')))
lib.setCurrentEyeShaderName(b'single_dimension_fast'); er.setRenderSize(lib,1000,1)
for S in (1,64,1000):
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    for _ in range(50): lib.renderFrame()
    host=[];dev=[]
    t0=time.perf_counter()
    for _ in range(500):
        host.append(lib.renderFrame()); dev.append(lib.crGetLastTraceMs())
    wall=(time.perf_counter()-t0)/500*1e3
    print(f"S={S}: renderFrame return {np.mean(host)*1e3:.1f} us, K1+K1b events {np.mean(dev)*1e3:.1f} us, python wall per frame {wall*1e3:.1f} us")
