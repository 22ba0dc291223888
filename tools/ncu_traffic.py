"""DRAM traffic per launch of the SPADE GEMM kernels at the bench shapes, from an `ncu --set full` capture of
    ncu --set full --clock-control none -k regex:'conv3x3_tc|wgrad3x3_tc' -c 12 -o <rep> python tools/ncu_targets.py spade 2 <C> 512 256 <images>
(two SPADE forward+backward passes; the second, warm one is used).  Launch order of one pass (ag2video_b200/spade.py):
conv epi1 (segmap -> actv), conv epi2 (actv -> gamma|beta + modulation), wgrad (gamma|beta), conv epi3 (input gradient,
gated), wgrad (shared), conv epi4 (accumulate into the segmap gradient).
    python tools/ncu_traffic.py <rep.ncu-rep> <C> <images> [more reps ...] > profiles/ncu_traffic.json
Keys match bench.py's (kernel, epilogue, resolution, Cin, Nout, images)."""
import csv
import json
import subprocess
import sys


def launches(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    ik, ir, iw, it = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'), h.index('gpu__time_duration.sum')
    units = rows[1]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    res = []
    for r in rows[2:]:
        res.append((r[ik], float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0), float(r[it]), units[it]))
    return res


def main():
    args = sys.argv[1:]
    result = {'source': [], 'launches': {}, 'detail': {}}
    for i in range(0, len(args), 3):
        rep, C, images = args[i], int(args[i + 1]), int(args[i + 2])
        ls = launches(rep)
        assert len(ls) >= 12, 'expected two passes of 6 launches, got %d' % len(ls)
        last = ls[-6:]
        keys = ['conv3x3 epi1 256 512 128 %d' % images, 'conv3x3 epi2 256 128 %d %d' % (2 * C, images),
                'wgrad3x3 wgrad 256 128 %d %d' % (2 * C, images), 'conv3x3 epi3 256 %d 128 %d' % (2 * C, images),
                'wgrad3x3 wgrad 256 512 128 %d' % images, 'conv3x3 epi4 256 128 512 %d' % images]
        order = ['conv3x3_tc', 'conv3x3_tc', 'wgrad3x3_tc', 'conv3x3_tc', 'wgrad3x3_tc', 'conv3x3_tc']
        for (name, nbytes, t, unit), key, want in zip(last, keys, order):
            assert want in name, (name, want)
            result['launches'][key] = nbytes
            result['detail'][key] = {'kernel': name.split('(')[0], 'duration': t, 'duration_unit': unit}
        result['source'].append(rep.replace('gpurun_out/', 'profiles/ (from gpurun_out/) '))
    result['source'] = 'ncu --set full --clock-control none of tools/ncu_targets.py spade: ' + ', '.join(result['source'])
    print(json.dumps(result, indent=1))


if __name__ == '__main__':
    main()
