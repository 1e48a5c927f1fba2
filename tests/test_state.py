"""State blobs: get/setStateInformation of the reference (e.g. JuicySaturator/PluginProcessor.cpp:117-131)
through the C ABI.  Host logic only -- no GPU.  JUCE is not available offline, so the blob format follows JUCE's
documentation (copyXmlToBinary: magic 0x21324356, length, single-line XML of the APVTS "PARAMS" tree, NUL) rather than a
golden file: parity of the FORMAT is unpinned; what is pinned here is the behaviour (round trip, replaceState semantics,
range handling identical to host automation, tolerance to real-world XML layout)."""
import struct

import pytest

from cases import PLUGINS


def blob_of(xml_text):
    data = xml_text.encode("utf-8")
    return struct.pack("<II", 0x21324356, len(data)) + data + b"\0"


@pytest.mark.parametrize("plugin", PLUGINS)
def test_state_round_trip(plugin, jb):
    a = jb.BatchProcessor(plugin, 4, device=-1)
    info = [p for p in a.parameterInfo() if not p["is_output"]]
    for i, p in enumerate(info):                       # move every parameter off its default
        a.setValueNotifyingHost(p["id"], 0.15 + 0.7 * ((i * 37) % 10) / 10.0)
    blob = a.getStateInformation()
    magic, length = struct.unpack("<II", blob[:8])
    assert magic == 0x21324356 and length == len(blob) - 9 and blob[-1] == 0
    text = blob[8:-1].decode("utf-8")
    assert text.startswith("<?xml") and "<PARAMS>" in text and text.count("<PARAM id=") == len(a.parameterInfo())
    b = jb.BatchProcessor(plugin, 4, device=-1)
    b.setStateInformation(blob)
    for p in a.parameterInfo():
        assert b.getRawParameterValue(p["id"]) == a.getRawParameterValue(p["id"]), p["id"]
    a.close()
    b.close()


def test_state_import_follows_replace_state(jb):
    eng = jb.BatchProcessor("JuicySaturator", 6, device=-1)
    ref = jb.BatchProcessor("JuicySaturator", 1, device=-1)
    eng.setParameter("mix", 0.25)                      # not in the blob below: goes back to its default
    default_mix = ref.getRawParameterValue("mix")
    # a blob as a DAW session might hold it: prolog, line breaks, value before id, single quotes, an unknown child,
    # an out-of-range value
    xml = """<?xml version="1.0" encoding="UTF-8"?>

<PARAMS>
  <PARAM id="drive" value="17.5"/>
  <PARAM value='0.33' id='asymmetry'/>
  <PARAM id="tone" value="7.0"/>
  <PARAM id="no_such_parameter" value="1.0"/>
</PARAMS>
"""
    eng.setStateInformation(blob_of(xml))
    ref.setParameter("drive", 17.5)
    ref.setParameter("asymmetry", 0.33)
    ref.setParameter("tone", 7.0)                      # clamped by the range, exactly like host automation
    for pid in ("drive", "asymmetry", "tone"):
        assert eng.getRawParameterValue(pid) == ref.getRawParameterValue(pid), pid
    assert eng.getRawParameterValue("mix") == default_mix
    # per clip range: clips 2..3 take another session's state, the rest keep theirs
    eng.setStateInformation(blob_of('<PARAMS><PARAM id="drive" value="3.0"/></PARAMS>'), first_clip=2, n_clips=2)
    assert eng.numParameterSets() == 2
    assert eng.getParameterClip("drive", 2) == pytest.approx(3.0)
    assert eng.getParameterClip("drive", 1) == ref.getRawParameterValue("drive")
    clip_blob = eng.getStateInformation(clip=3)
    assert b'id="drive" value="3.0"' in clip_blob
    ref.close()
    eng.close()


def test_state_errors(jb):
    eng = jb.BatchProcessor("JuicyWidth", 2, device=-1)
    before = eng.getRawParameterValue("width")
    with pytest.raises(jb.JuicyBatchError):
        eng.setStateInformation(b"not a blob at all")
    with pytest.raises(jb.JuicyBatchError):
        eng.setStateInformation(blob_of("<OTHER><PARAM id=\"width\" value=\"0.1\"/></OTHER>"))   # another tree type
    with pytest.raises(jb.JuicyBatchError):
        eng.setStateInformation(blob_of("<PARAMS/>"), first_clip=1, n_clips=5)
    assert eng.getRawParameterValue("width") == before
    eng.close()
