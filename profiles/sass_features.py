"""SASS listing summary of the built library: instructions per kernel and the Blackwell / Hopper data-movement instructions in them.
Usage: python profiles/sass_features.py h264-scroll-encoder_b200/libh264scroll_b200.so > profiles/<round>_sass_features.txt"""
import subprocess, sys, re, collections
lib = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
feat = re.compile(r"\b(UTMALDG\S*|UTMASTG\S*|UBLKCP\S*|UBLKPF\S*|UMEMSETS\S*|SYNCS\S*|LDGSTS\S*)")
fn = None; count = collections.OrderedDict(); feats = {}; first = {}
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); count[fn] = 0; feats[fn] = collections.Counter(); first[fn] = {}; continue
    if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        count[fn] += 1
        m = feat.search(line)
        if m:
            k = m.group(1); feats[fn][k] += 1
            first[fn].setdefault(k, re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", line).rstrip(" ;").strip())
names = subprocess.run(["c++filt"] + list(count), capture_output=True, text=True).stdout.split("\n")
print("cuobjdump -sass %s (sm_100a), round-2 final build: instructions per kernel and the" % lib)
print("Blackwell / Hopper data-movement instructions in them (UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, UBLKPF = bulk L2")
print("prefetch, UMEMSETS = st.bulk, SYNCS = mbarrier operations), with the first occurrence of each as it stands in the listing\n")
for (f, c), n in sorted(zip(count.items(), names), key=lambda t: -t[0][1]):
    print("%-94s %6d instructions" % (n[:94], c))
    if feats[f]:
        print("    " + "  ".join("%s x%d" % kv for kv in sorted(feats[f].items())))
        for k in sorted(first[f]): print("        " + first[f][k])
