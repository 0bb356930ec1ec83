#!/usr/bin/env python3
"""loudgain's scan for WAV files on the B200 library, printed like `loudgain -O`
(reference: /root/reference/src/loudgain.c:299-379 scan loop and clipping
prevention, :586-612 the tab-separated rows).  Decoding other containers
(FFmpeg) and writing tags (TagLib) stay outside this repository.

    python tools/loudgain_wav.py [-a] [-k | -K dBTP] [-d pregain] [--threads N] a.wav b.wav ...
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class ClipInfo(C.Structure):
    _fields_ = [("will_clip", C.c_int), ("track_clipped", C.c_int), ("album_clipped", C.c_int),
                ("album_would_clip", C.c_int), ("track_new_peak", C.c_double),
                ("album_new_peak", C.c_double)]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-a", "--album", action="store_true", help="also compute album gain / peak / range")
    ap.add_argument("-k", "--noclip", action="store_true", help="lower the gain to stay below -1 dBTP")
    ap.add_argument("-K", "--maxtpl", type=float, default=None, help="like -k with this limit (dBTP)")
    ap.add_argument("-d", "--pregain", type=float, default=0.0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("files", nargs="+")
    args = ap.parse_args()

    from loudgain_b200 import engine, wavio
    tracks = [wavio.read_wav(f) for f in args.files]
    results = engine.scan_host(tracks, chunk_frames=4096, do_album=args.album, pre_gain=args.pregain,
                               threads=max(1, min(args.threads, len(tracks))))
    L = engine._bind()
    L.lgb_clip_prevention.argtypes = [C.POINTER(engine.ScanResult), C.c_int, C.c_int, C.c_double,
                                      C.POINTER(ClipInfo)]
    L.lgb_format_tab_row.argtypes = [C.c_char_p, C.POINTER(engine.ScanResult), C.POINTER(ClipInfo), C.c_int,
                                     C.c_char_p, C.c_char_p, C.c_size_t]
    L.lgb_format_tab_row.restype = C.c_size_t
    limit = args.maxtpl if args.maxtpl is not None else -1.0
    prevent = args.noclip or args.maxtpl is not None
    print("File\tLoudness\tRange\tTrue_Peak\tTrue_Peak_dBTP\tReference\tWill_clip\tClip_prevent\tGain\t"
          "New_Peak\tNew_Peak_dBTP")
    buf = C.create_string_buffer(1024)
    for i, (name, r) in enumerate(zip(args.files, results)):
        info = ClipInfo()
        L.lgb_clip_prevention(C.byref(r), int(args.album), int(prevent), limit, C.byref(info))
        L.lgb_format_tab_row(name.encode(), C.byref(r), C.byref(info), 0, b"dB", buf, len(buf))
        sys.stdout.write(buf.value.decode())
        if args.album and i == len(results) - 1:
            L.lgb_format_tab_row(b"Album", C.byref(r), C.byref(info), 1, b"dB", buf, len(buf))
            sys.stdout.write(buf.value.decode())


if __name__ == "__main__":
    main()
