import sys, os
sys.path.insert(0,'/root/repo/compound-ray_b200'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/benchmarks')
import eye_renderer as er, speed_test
data=speed_test.fixtures()
lib=er.load_library(device=0); lib.setVerbosity(False)
lib.loadGlTFscene(os.path.join(data,'data','natural-standin-sky.gltf').encode())
er.gotoFirstCompoundEye(lib)
er.setOmmatidiaFromOmmatidiumList(lib, er.readEyeFile(os.path.join(data,'data','eyes','1000-equidistant.eye')))
lib.setCurrentEyeShaderName(b'single_dimension_fast'); er.setRenderSize(lib,1000,1)
for S in (1,64):
    lib.setCurrentEyeSamplesPerOmmatidium(S)
    for _ in range(6): lib.renderFrame()
