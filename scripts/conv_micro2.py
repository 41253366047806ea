import sys, os
sys.path.insert(0, os.getcwd())
sys.argv=['x']
exec(open('scripts/conv_micro.py').read().split("print(\"--- 1x1 tiled vs im2col A loads\")")[0])
print("--- stage depth, pair 256x256 tile")
for stg in (2,3,4,5):
    bench(64, 26, 256, 512, 3, 1, pair=2, bn=256, stages=stg)
for stg in (2,3,4,5):
    bench(64, 13, 512, 1024, 3, 1, pair=2, bn=256, stages=stg)
print("--- 1x1")
for stg in (3,5,7):
    bench(64, 52, 256, 128, 1, 1, pair=2, stages=stg)
    bench(64, 26, 512, 256, 1, 1, pair=2, stages=stg)
for bn in (128, 256):
    bench(64, 26, 512, 256, 1, 1, pair=2, bn=bn)
    bench(64, 13, 1024, 512, 1, 1, pair=2, bn=bn)
    bench(64, 13, 512, 1024, 3, 1, pair=2, bn=bn)
for pair in (1,2):
    bench(64, 52, 256, 128, 1, 1, pair=pair)
