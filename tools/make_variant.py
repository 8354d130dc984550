"""Experiment builds: python tools/make_variant.py <name> [-DFLAG=…] -> restir-vulkan_b200/variants/lib_<name>.so (loaded through RESTIR_B200_LIB)."""
import sys, os, importlib.util
spec = importlib.util.spec_from_file_location("_restir_build", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "restir-vulkan_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
name = sys.argv[1]; defs = sys.argv[2:]
out = os.path.join(b.HERE, 'variants', 'lib_%s.so' % name)
os.makedirs(os.path.dirname(out), exist_ok=True)
b.build_variant(out, defs, verbose=False)
print(out)
